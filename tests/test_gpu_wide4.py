"""Tensor-core forward of the 4-layer ODE_01 DE_Func at the scripts' argparse default `--hidden 128`
(neural_00_ODE_01_no_encode.py:61-68, :245-247; `psn_wide4_fwd_kernel`, reached by impl = wide and by impl = auto): against the CPU
oracle at rtol=1e-5 / atol=1e-6 (all three schemes, narrow / odd state and input widths, ragged batch, events, batch-major
storage views), against the CUDA-core generic kernel over more than one wave of CTAs, and repeated runs for determinism."""
import os

import pytest
import torch

from helpers import ATOL, RTOL, tol_report

pytestmark = pytest.mark.gpu
H = 128


def _params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]


def _problem(B, N, seed, X=16, Z=2, events=0, scale=0.1, hidden=H):
    from py_psnode_b200 import DE_Func
    torch.manual_seed(seed)
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=hidden)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, X) * scale
    z = torch.randn(T, B, Z) * scale
    a0 = torch.cat((x[0], z[0]), dim=-1)
    ev = None
    if events:
        steps = [N // 3, (2 * N) // 3][:events]
        event_t = torch.stack([t[s, :, 0] for s in steps], dim=1).view(B, events, 1).clone()
        z_jump = torch.randn(B, events, Z) * scale
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
def test_wide4_forward_vs_oracle(native_lib, solver):
    de, t, x, z, a0, ev = _problem(B=48, N=24, seed=61)
    want = _oracle(solver, de, t, x, z, a0, ev)
    got, kern = _run(solver, de, t, x, z, a0, ev, "wide")
    assert kern.startswith("psn_wide4_fwd_kernel"), kern
    assert torch.equal(got[0], x[0])
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want)


@pytest.mark.parametrize("X,Z", [(5, 3), (16, 8), (1, 1), (12, 5)])
def test_wide4_forward_narrow_widths_ragged_batch_events_batch_major(native_lib, X, Z):
    """State narrower than the 16 padded rows, held inputs up to the limit of 8, B = 37 (3 groups, the last with 5 live
    trajectories), two events, series passed as permuted views of (B,T,.) storage."""
    de, t, x, z, a0, ev = _problem(B=37, N=30, seed=63 + X, X=X, Z=Z, events=2)
    want = _oracle("rk4", de, t, x, z, a0, ev)
    got, kern = _run("rk4", de, t, x, z, a0, ev, "wide", batch_major=True)
    assert kern.startswith("psn_wide4_fwd_kernel"), kern
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want)
    no_ev = _oracle("rk4", de, t, x, z, a0, None)
    assert not torch.allclose(no_ev, want, rtol=1e-3, atol=1e-4), "the events must change the trajectory"


@pytest.mark.parametrize("hidden,impl", [(96, "auto"), (100, "auto"), (72, "wide"), (40, "wide"), (64, "wide")])
def test_wide4_forward_padded_hidden_widths(native_lib, hidden, impl):
    """Hidden widths below 128 run zero-padded to 128 neurons (exact: a padded neuron has zero weights and bias, ELU(0) = 0);
    impl = auto sends 64 < H <= 128 here (H = 64 has its own kernel, narrower nets stay on the CUDA cores), impl = wide any H <= 128."""
    de, t, x, z, a0, ev = _problem(B=21, N=20, seed=70 + hidden, events=1, hidden=hidden)
    want = _oracle("rk4", de, t, x, z, a0, ev)
    got, kern = _run("rk4", de, t, x, z, a0, ev, impl)
    if impl == "wide" or not os.environ.get("PSNODE_WIDE4", "1").startswith("0"):
        assert kern.startswith("psn_wide4_fwd_kernel"), kern
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want)


def test_wide4_forward_vs_generic_many_ctas(native_lib):
    """B = 5000 -> 313 groups on 157 CTAs (more than one wave): every CTA / group position against the CUDA-core kernel;
    120 steps so the one-step-ahead staging buffers wrap many times; a second run must be bit-identical."""
    de, t, x, z, a0, ev = _problem(B=5000, N=120, seed=64, events=1)
    ref, k0 = _run("rk4", de, t, x, z, a0, ev, "generic")
    got, k1 = _run("rk4", de, t, x, z, a0, ev, "wide")
    assert k0.startswith("psn_generic_fwd_kernel") and k1.startswith("psn_wide4_fwd_kernel"), (k0, k1)
    assert torch.allclose(got, ref, rtol=RTOL, atol=ATOL), tol_report(got, ref)
    again, _ = _run("rk4", de, t, x, z, a0, ev, "wide")
    assert torch.equal(again, got), "the kernel must be deterministic"


def test_wide4_is_what_auto_picks_and_training_still_matches(native_lib):
    """impl = auto reaches the kernel (unless PSNODE_WIDE4=0), and a training step through it with the switches at their defaults
    (tensor-core forward with tape + tensor-core reverse sweep) gives the parameter gradients of float64 autograd through the oracle."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import RK4, _native
    dev = "cuda:0"
    de, t, x, z, a0, ev = _problem(B=40, N=12, seed=65)
    got, kern = _run("rk4", de, t, x, z, a0, ev, "auto")
    expect = "psn_generic_fwd_kernel" if os.environ.get("PSNODE_WIDE4", "1").startswith("0") else "psn_wide4_fwd_kernel"
    assert kern.startswith(expect), kern
    # float64 reference gradients
    p64 = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.cpu().x_dot)]
    want = O.integrate_ode("rk4", p64, t.double(), x.double(), z.double(), a0.double())
    want.square().sum().backward()
    ref = [g for W, b in p64 for g in (W.grad, b.grad)]
    de = de.to(dev)
    for p in de.parameters():
        p.grad = None
    out = RK4().integrate_ODE(x_func=de, t=t.to(dev), x=x.to(dev), z=z.to(dev), all_initial=a0.to(dev))
    out.square().sum().backward()
    if os.environ.get("PSNODE_WIDE4", "1")[:1] != "0" and os.environ.get("PSNODE_WIDE4_BWD", "1")[:1] != "0":
        assert _native.last_kernel().startswith("psn_wide4_assemble_kernel"), _native.last_kernel()
    for p, g in zip(de.parameters(), ref):
        err = (p.grad.double().cpu() - g).abs().max().item()
        assert err <= 2e-5 * max(g.abs().max().item(), 1e-3), (tuple(p.shape), err, g.abs().max().item())


# ---- tensor-core reverse sweep of the same shape (psn_wide4_bwd_kernel + psn_wide_grad_kernel + psn_wide4_assemble_kernel) -------------
@pytest.mark.parametrize("solver,X,Z,hidden,B,events", [("rk4", 16, 2, 128, 40, 1), ("euler", 16, 2, 128, 16, 0), ("midpoint", 5, 3, 96, 37, 2),
                                                        ("rk4", 12, 8, 128, 21, 0), ("rk4", 1, 1, 72, 5, 1)])
def test_wide4_gradients_vs_fp64_autograd(native_lib, monkeypatch, solver, X, Z, hidden, B, events):
    """Discrete adjoint on the tensor cores for the 4-layer net: every parameter gradient, dL/dx[0] and dL/d all_initial against float64
    autograd through the oracle (the switch is read per call: PSNODE_WIDE4_BWD)."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import Euler, Midpoint, ODE_Event, RK4, _native
    monkeypatch.setenv("PSNODE_WIDE4_BWD", "1")
    dev = "cuda:0"
    N = 14
    de, t, x, z, a0, ev = _problem(B=B, N=N, seed=80 + X + B, X=X, Z=Z, events=events, hidden=hidden)
    T = N + 1
    w = torch.randn(T, B, X) * 0.1
    p64 = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    x64 = x.double().requires_grad_(True)
    a064 = a0.double().requires_grad_(True)
    args = (ev[0].double(), ev[1].double()) if ev else ()
    sol64 = O.integrate_ode(solver, p64, t.double(), x64, z.double(), a064, *args)
    (sol64 * w.double()).sum().backward()
    de_d = de.to(dev)
    xd = x.to(dev).requires_grad_(True)
    a0d = a0.to(dev).requires_grad_(True)
    kw = {}
    if ev:
        e = ODE_Event()
        e.set_event(t=ev[0].to(dev), z=ev[1].to(dev))
        kw = dict(event_fn=e.event_fn, jump_change_fn=e.jump_change_fn)
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[solver]
    sol = S(impl="wide").integrate_ODE(x_func=de_d, t=t.to(dev), x=xd, z=z.to(dev), all_initial=a0d, **kw)
    assert _native.last_kernel().startswith("psn_wide4_fwd_kernel") and "tape" in _native.last_kernel(), _native.last_kernel()
    assert torch.allclose(sol.detach().cpu(), sol64.detach().float(), rtol=RTOL, atol=ATOL)
    (sol * w.to(dev)).sum().backward()
    assert _native.last_kernel().startswith("psn_wide4_assemble_kernel"), _native.last_kernel()
    lin = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    pairs = [(f"{n}{k + 1}", getattr(lin[k], a).grad, p64[k][i].grad) for k in range(4) for n, a, i in (("W", "weight", 0), ("b", "bias", 1))]
    pairs += [("x", xd.grad, x64.grad), ("all_initial", a0d.grad, a064.grad)]
    bad = []
    for name, g, g64 in pairs:
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        print(f"grad {name}: max err {err:.3e} scale {scale:.3e} rel {err / max(scale, 1e-30):.2e}")
        if not err <= 2e-5 * scale + 1e-7:
            bad.append((name, err, scale))
    assert not bad, bad


def test_wide4_fused_loss_sweep_matches_generic_sweep_many_ctas(native_lib, monkeypatch):
    """B = 2500 (157 groups, 79 CTAs of the sweep; 148 CTAs of the block-GEMM pass), 40 RK4 steps, one event, masked-MSE loss fused into the
    sweep: parameter gradients against the generic recomputing sweep of the same problem, twice for determinism."""
    from py_psnode_b200 import ODE_Event, RK4, _native
    dev = "cuda:0"
    B, N = 2500, 40
    de, t, x, z, a0, ev = _problem(B=B, N=N, seed=91, events=1)
    T = N + 1
    target = (torch.randn(T, B, 16) * 0.1).to(dev)
    mask = (torch.rand(T, B, 1) > 0.3).float().to(dev)
    de_d = de.to(dev)
    e = ODE_Event()
    e.set_event(t=ev[0].to(dev), z=ev[1].to(dev))

    def grads(flag, impl):
        monkeypatch.setenv("PSNODE_WIDE4_BWD", flag)
        for p in de_d.parameters():
            p.grad = None
        num, _ = RK4(impl=impl).integrate_ODE_loss(x_func=de_d, t=t.to(dev), x=x.to(dev), z=z.to(dev), all_initial=a0.to(dev), target=target,
                                                   mask=mask, event_fn=e.event_fn, jump_change_fn=e.jump_change_fn)
        (num / mask.sum()).backward()
        return [p.grad.clone() for p in de_d.parameters()], _native.last_kernel()

    ref, k0 = grads("0", "wide")
    got, k1 = grads("1", "wide")
    again, _ = grads("1", "auto")
    assert k0.startswith("psn_grad_reduce_kernel") or "generic" in k0, k0
    assert k1.startswith("psn_wide4_assemble_kernel"), k1
    for p, a, b, c in zip(de_d.parameters(), ref, got, again):
        scale = float(a.abs().max())
        err = float((a - b).abs().max())
        assert err <= 2e-5 * scale + 1e-9, (tuple(p.shape), err, scale)
        assert torch.equal(b, c), "the sweep must be deterministic"
