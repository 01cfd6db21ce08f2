"""Per-layer tcgen05 GEMM path (impl = layer) for the latent nets too wide for one SM -- DAE_02 / ODE_02 with
X = Z (= V = I) = hidden = 128 or 256 (BASELINE configs[4]; neural_01_DAE_02_direct_encode.py:70-100,137-147): forward against
the CPU oracle at rtol=1e-5 / atol=1e-6 (all schemes, events, ragged batch, batch-major views) and against the CUDA-core
generic kernel over many CTAs."""
import pytest
import torch

from helpers import ATOL, RTOL, tol_report

pytestmark = pytest.mark.gpu


def _params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]


def _dae_problem(B, N, H, seed, events=0, scale=0.05):
    from py_psnode_b200 import AE_Func, DE_Func
    torch.manual_seed(seed)
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: torch.randn(T, B, H) * scale
    d = dict(t=t, x=mk(), z=mk(), v=mk(), i=mk(), x_init=torch.randn(B, H) * scale)
    d["a0"] = torch.cat((d["x_init"], d["z"][0], d["v"][0], d["i"][0]), dim=-1)
    ev = None
    if events:
        steps = [N // 3, (2 * N) // 3][:events]
        event_t = torch.stack([t[s, :, 0] for s in steps], dim=1).view(B, events, 1).clone()
        ev = (event_t, torch.randn(B, events, H) * scale, torch.randn(B, events, H) * scale)
    return de, ae, d, ev


def _run_dae(solver_name, de, ae, d, ev, impl, batch_major=False, dev="cuda:0"):
    from py_psnode_b200 import DAE_Event, Euler, Midpoint, RK4, _native
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[solver_name]
    kw = {}
    if ev is not None:
        e = DAE_Event()
        e.set_event(t=ev[0].to(dev), z=ev[1].to(dev), v=ev[2].to(dev))
        kw = dict(event_fn=e.event_fn, jump_change_fn=e.jump_change_fn)
    mv = (lambda q: q.permute(1, 0, 2).contiguous().to(dev).permute(1, 0, 2)) if batch_major else (lambda q: q.to(dev))
    with torch.no_grad():
        gx, gi = S(impl=impl).integrate_DAE(x_init=d["x_init"].to(dev), x_func=de.to(dev), i_func=ae.to(dev), t=mv(d["t"]), x=mv(d["x"]),
                                            z=mv(d["z"]), v=mv(d["v"]), i=mv(d["i"]), all_initial=d["a0"].to(dev), **kw)
    return gx.cpu(), gi.cpu(), _native.last_kernel()


def _oracle_dae(solver_name, de, ae, d, ev):
    from oracle import psnode_oracle as O
    de, ae = de.cpu(), ae.cpu()
    args = (solver_name, _params(de.x_dot), _params(ae.i_calculator), d["x_init"], d["t"], d["x"], d["z"], d["v"], d["i"], d["a0"])
    return O.integrate_dae(*args) if ev is None else O.integrate_dae(*args, ev[0], ev[1], ev[2])


@pytest.mark.parametrize("H,solver,events", [(256, "rk4", 1), (256, "euler", 0), (128, "midpoint", 2), (128, "rk4", 0)])
def test_layer_dae_forward_vs_oracle(native_lib, H, solver, events):
    de, ae, d, ev = _dae_problem(B=24, N=10, H=H, seed=43 + H, events=events)
    wx, wi = _oracle_dae(solver, de, ae, d, ev)
    gx, gi, kern = _run_dae(solver, de, ae, d, ev, "auto")
    assert kern.startswith("psn_lg_gemm_kernel"), kern
    assert torch.equal(gx[0], d["x_init"])
    assert torch.allclose(gx, wx, rtol=RTOL, atol=ATOL), "x: " + tol_report(gx, wx)
    assert torch.allclose(gi, wi, rtol=RTOL, atol=ATOL), "i: " + tol_report(gi, wi)


def test_layer_dae_ragged_batch_major_many_ctas(native_lib):
    """B = 300 (3 n-tiles, the last with 44 live rows) x 2 M-blocks, 25 steps, two events, permuted (B,T,.) storage: against the
    oracle on sampled rows and against the generic CUDA-core kernel over the whole batch; deterministic."""
    de, ae, d, ev = _dae_problem(B=300, N=25, H=256, seed=47, events=2)
    gx, gi, kern = _run_dae("rk4", de, ae, d, ev, "layer", batch_major=True)
    assert kern.startswith("psn_lg_gemm_kernel"), kern
    rx, ri, k0 = _run_dae("rk4", de, ae, d, ev, "generic")
    assert k0.startswith("psn_generic_fwd_kernel"), k0
    assert torch.allclose(gx, rx, rtol=RTOL, atol=ATOL), "x: " + tol_report(gx, rx)
    assert torch.allclose(gi, ri, rtol=RTOL, atol=ATOL), "i: " + tol_report(gi, ri)
    ax, ai, _ = _run_dae("rk4", de, ae, d, ev, "layer", batch_major=True)
    assert torch.equal(ax, gx) and torch.equal(ai, gi)
    rows = [0, 127, 128, 255, 256, 299]
    sub = {k: (v[:, rows] if v.dim() == 3 else v[rows]) for k, v in d.items()}
    sev = (ev[0][rows], ev[1][rows], ev[2][rows])
    # the event predicate looks at sample 0 of the batch it is given (neural_base.py:54): row 0 is in the sample
    wx, wi = _oracle_dae("rk4", de, ae, sub, sev)
    assert torch.allclose(gx[:, rows], wx, rtol=RTOL, atol=ATOL), "x rows: " + tol_report(gx[:, rows], wx)
    assert torch.allclose(gi[:, rows], wi, rtol=RTOL, atol=ATOL), "i rows: " + tol_report(gi[:, rows], wi)


@pytest.mark.parametrize("solver", ["rk4", "euler"])
def test_layer_ode_h256_vs_oracle(native_lib, solver):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, Euler, ODE_Event, RK4, _native
    torch.manual_seed(51)
    dev = "cuda:0"
    B, N, H = 40, 12, 256
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x, z = torch.randn(T, B, H) * 0.05, torch.randn(T, B, H) * 0.05
    a0 = torch.cat((x[0], z[0]), dim=-1)
    event_t = t[N // 2].view(B, 1, 1).clone()
    z_jump = torch.randn(B, 1, H) * 0.05
    want = O.integrate_ode(solver, _params(de.x_dot), t, x, z, a0, event_t, z_jump)
    ev = ODE_Event()
    ev.set_event(t=event_t.to(dev), z=z_jump.to(dev))
    S = {"euler": Euler, "rk4": RK4}[solver]
    with torch.no_grad():
        got = S().integrate_ODE(x_func=de.to(dev), t=t.to(dev), x=x.to(dev), z=z.to(dev), all_initial=a0.to(dev), event_fn=ev.event_fn,
                                jump_change_fn=ev.jump_change_fn).cpu()
    assert _native.last_kernel().startswith("psn_lg_gemm_kernel"), _native.last_kernel()
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want)
