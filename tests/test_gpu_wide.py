"""Tensor-core path of the latent `*_02_direct_encode` nets (impl = wide; BASELINE configs[3]: X = Z = H = 128, 2-layer DE_Func,
neural_00_ODE_02_direct_encode.py:49-57,70): forward against the CPU oracle at rtol=1e-5 / atol=1e-6 (all three schemes,
ragged batch, events, batch-major storage views), against the CUDA-core generic kernel over many CTAs, the TMA-staged
projection kernel against its CUDA-core twin, and gradients (parameters, x[0], all_initial, the latent input series and jump
tensors the encoders need -- SURVEY 3.3) against float64 autograd through the oracle."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ATOL, RTOL, tol_report

pytestmark = pytest.mark.gpu
H = 128


def _params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]


def _problem(B, N, seed, events=0, scale=0.1):
    from py_psnode_b200 import DE_Func
    torch.manual_seed(seed)
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, H) * scale
    z = torch.randn(T, B, H) * scale
    a0 = torch.cat((x[0], z[0]), dim=-1)
    ev = None
    if events:
        steps = [N // 3, (2 * N) // 3][:events]
        event_t = torch.stack([t[s, :, 0] for s in steps], dim=1).view(B, events, 1).clone()
        z_jump = torch.randn(B, events, H) * scale
        ev = (event_t, z_jump)
    return de, t, x, z, a0, ev


def _run(solver_name, de, t, x, z, a0, ev, impl, dev="cuda:0", batch_major=False):
    from py_psnode_b200 import Euler, Midpoint, ODE_Event, RK4, _native
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[solver_name]
    kw = {}
    if ev is not None:
        e = ODE_Event()
        e.set_event(t=ev[0].to(dev), z=ev[1].to(dev))
        kw = dict(event_fn=e.event_fn, jump_change_fn=e.jump_change_fn)
    mv = (lambda q: q.permute(1, 0, 2).contiguous().to(dev).permute(1, 0, 2)) if batch_major else (lambda q: q.to(dev))
    with torch.no_grad():
        got = S(impl=impl).integrate_ODE(x_func=de.to(dev), t=mv(t), x=mv(x), z=mv(z), all_initial=a0.to(dev), **kw)
    return got.cpu(), _native.last_kernel()


def _oracle(solver_name, de, t, x, z, a0, ev):
    from oracle import psnode_oracle as O
    de = de.cpu()
    if ev is None:
        return O.integrate_ode(solver_name, _params(de.x_dot), t, x, z, a0)
    return O.integrate_ode(solver_name, _params(de.x_dot), t, x, z, a0, ev[0], ev[1])


@pytest.mark.parametrize("solver", ["euler", "midpoint", "rk4"])
def test_wide_forward_vs_oracle(native_lib, solver):
    de, t, x, z, a0, ev = _problem(B=48, N=24, seed=41)
    want = _oracle(solver, de, t, x, z, a0, ev)
    got, kern = _run(solver, de, t, x, z, a0, ev, "auto")
    assert kern.startswith("psn_wide_fwd_kernel"), kern
    assert torch.equal(got[0], x[0])
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want)


def test_wide_forward_ragged_batch_events_batch_major(native_lib):
    """B = 37 (3 groups, the last with 5 live trajectories), two events, series passed as permuted views of (B,T,.) storage."""
    de, t, x, z, a0, ev = _problem(B=37, N=30, seed=43, events=2)
    want = _oracle("rk4", de, t, x, z, a0, ev)
    got, kern = _run("rk4", de, t, x, z, a0, ev, "wide", batch_major=True)
    assert kern.startswith("psn_wide_fwd_kernel"), kern
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want)
    no_ev = _oracle("rk4", de, t, x, z, a0, None)
    assert not torch.allclose(no_ev, want, rtol=1e-3, atol=1e-4), "the events must change the trajectory"


def test_wide_forward_vs_generic_many_ctas(native_lib):
    """B = 5000 -> 313 groups on 157 CTAs (more than one wave): every CTA / group / TMA tile position against the CUDA-core
    kernel; 120 steps so the prefetch buffers wrap many times."""
    de, t, x, z, a0, ev = _problem(B=5000, N=120, seed=44, events=1)
    ref, k0 = _run("rk4", de, t, x, z, a0, ev, "generic")
    got, k1 = _run("rk4", de, t, x, z, a0, ev, "wide")
    assert k0.startswith("psn_generic_fwd_kernel") and k1.startswith("psn_wide_fwd_kernel")
    assert torch.allclose(got, ref, rtol=RTOL, atol=ATOL), tol_report(got, ref)
    again, _ = _run("rk4", de, t, x, z, a0, ev, "wide")
    assert torch.equal(again, got), "the kernel must be deterministic"


def test_wide_projection_tma_vs_cuda_core(native_lib):
    """The TMA + tcgen05 projection (UTMALDG-staged series tiles) against its fp32 CUDA-core twin (PSNODE_WIDE_PROJ=simple),
    in a second process because the switch is read once per process."""
    de, t, x, z, a0, ev = _problem(B=200, N=16, seed=45, events=1)
    got, _ = _run("euler", de, t, x, z, a0, ev, "wide")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_wide_simple.pt")
    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import test_gpu_wide as W\n"
        "de, t, x, z, a0, ev = W._problem(B=200, N=16, seed=45, events=1)\n"
        "got, k = W._run('euler', de, t, x, z, a0, ev, 'wide')\n"
        "torch.save(got, %r)\n" % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))), path))
    env = dict(os.environ, PSNODE_WIDE_PROJ="simple")
    subprocess.run([sys.executable, "-c", code], check=True, env=env)
    simple = torch.load(path)
    os.remove(path)
    assert torch.allclose(got, simple, rtol=2e-6, atol=2e-7), tol_report(got, simple)


@pytest.mark.parametrize("solver,events", [("rk4", 1), ("euler", 0), ("midpoint", 2)])
def test_wide_gradients_vs_fp64_autograd(native_lib, solver, events):
    """Discrete adjoint on the tensor cores: every gradient sink of the encoded ODE model against float64 autograd."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import Euler, Midpoint, ODE_Event, RK4, _native
    dev = "cuda:0"
    B, N = 40, 14
    de, t, x, z, a0, ev = _problem(B=B, N=N, seed=46, events=events)
    T = N + 1
    w = torch.randn(T, B, H) * 0.1
    p64 = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    x64, z64 = x.double().requires_grad_(True), z.double().requires_grad_(True)
    a064 = torch.cat((x64[0], z64[0]), dim=-1)
    zj64 = ev[1].double().requires_grad_(True) if ev else None
    if ev:
        sol64 = O.integrate_ode(solver, p64, t.double(), x64, z64, a064, ev[0].double(), zj64)
    else:
        sol64 = O.integrate_ode(solver, p64, t.double(), x64, z64, a064)
    (sol64 * w.double()).sum().backward()
    de_d = de.to(dev)
    xd, zd = x.to(dev).requires_grad_(True), z.to(dev).requires_grad_(True)
    a0d = torch.cat((xd[0], zd[0]), dim=-1)
    kw = {}
    zjd = None
    if ev:
        zjd = ev[1].to(dev).requires_grad_(True)
        e = ODE_Event()
        e.set_event(t=ev[0].to(dev), z=zjd)
        kw = dict(event_fn=e.event_fn, jump_change_fn=e.jump_change_fn)
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[solver]
    sol = S().integrate_ODE(x_func=de_d, t=t.to(dev), x=xd, z=zd, all_initial=a0d, **kw)
    assert _native.last_kernel().startswith("psn_wide_fwd_kernel"), _native.last_kernel()
    (sol * w.to(dev)).sum().backward()
    assert _native.last_kernel().startswith("psn_wide"), _native.last_kernel()
    lin = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    pairs = [("W1", lin[0].weight.grad, p64[0][0].grad), ("b1", lin[0].bias.grad, p64[0][1].grad),
             ("W2", lin[1].weight.grad, p64[1][0].grad), ("b2", lin[1].bias.grad, p64[1][1].grad),
             ("x", xd.grad, x64.grad), ("z", zd.grad, z64.grad)]
    if ev:
        pairs.append(("z_jump", zjd.grad, zj64.grad))
    bad = []
    for name, g, g64 in pairs:
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        print(f"grad {name}: max err {err:.3e} scale {scale:.3e} rel {err / max(scale, 1e-30):.2e}")
        if not err <= 2e-5 * scale + 1e-7:
            bad.append((name, err, scale))
    assert not bad, bad
