"""All three schemes (Euler / Midpoint / RK4-3/8, neural_dae/my_fixed_grid.py:15-59) through the tensor-core kernels on a
ragged batch (B = 40: two and a half 16-trajectory groups) with an event: forward against the fp32 oracle at
rtol=1e-5 / atol=1e-6, and the tape-based reverse sweep against float64 autograd through the oracle."""
import pytest
import torch

from helpers import ATOL, RTOL, tol_report

pytestmark = pytest.mark.gpu


def _params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]


@pytest.mark.parametrize("X", [16, 6, 11])
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_ode_tc_forward_and_tape_gradients(native_lib, method, X, monkeypatch):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, Euler, Midpoint, ODE_Event, RK4, _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    torch.manual_seed(51)
    dev = "cuda:0"
    B, N, Z, H = 40, 30, 2, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H)
    t = (torch.arange(T, dtype=torch.float32) * 0.02).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, X) * 0.2
    z = torch.randn(T, B, Z) * 0.2
    w = torch.randn(T, B, X) * 0.1
    event_t = t[N // 3].view(B, 1, 1).clone()
    z_jump = torch.randn(B, 1, Z) * 0.2
    a0 = torch.cat((x[0], z[0]), dim=-1)
    want = O.integrate_ode(method, _params(de.x_dot), t, x, z, a0, event_t, z_jump)
    # float64 autograd through the oracle
    p64 = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    a064 = a0.double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    s64 = O.integrate_ode(method, p64, t.double(), x64, z.double(), a064, event_t.double(), z_jump.double())
    (s64 * w.double()).sum().backward()
    # CUDA
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[method]
    de_d = de.to(dev)
    ev = ODE_Event()
    ev.set_event(t=event_t.to(dev), z=z_jump.to(dev))
    xd = x.to(dev).requires_grad_(True)
    a0d = a0.to(dev).requires_grad_(True)
    sol = S().integrate_ODE(x_func=de_d, t=t.to(dev), x=xd, z=z.to(dev), all_initial=a0d, event_fn=ev.event_fn,
                            jump_change_fn=ev.jump_change_fn)
    assert _native.last_kernel().startswith("psn_tc8_ode_kernel")
    assert torch.allclose(sol.detach().cpu(), want, rtol=RTOL, atol=ATOL), tol_report(sol.detach().cpu(), want)
    (sol * w.to(dev)).sum().backward()
    assert _native.last_kernel() == "psn_tc_grad_reduce_kernel"
    lin = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    pairs = [(lin[k].weight.grad, p64[k][0].grad) for k in range(4)] + [(lin[k].bias.grad, p64[k][1].grad) for k in range(4)]
    pairs += [(a0d.grad, a064.grad), (xd.grad[0], x64.grad[0])]
    for k, (g, g64) in enumerate(pairs):
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 1e-5 * scale + 1e-7, f"{method} tensor {k}: err {err:.3e} scale {scale:.3e}"
    assert float(xd.grad[1:].abs().max()) == 0.0, "only x[0] (the initial state) receives a gradient without teacher forcing"


@pytest.mark.parametrize("X", [16, 5])
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_dae_tc_forward_all_methods(native_lib, method, X):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, Euler, Midpoint, RK4, _native
    torch.manual_seed(52)
    dev = "cuda:0"
    B, N, Z, V, I, H = 40, 30, 1, 2, 4, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z)
    t = (torch.arange(T, dtype=torch.float32) * 0.02).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda wd: torch.randn(T, B, wd) * 0.2
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    x_init = torch.randn(B, X) * 0.2
    a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
    event_t = torch.cat((t[1].view(B, 1, 1), t[N // 2].view(B, 1, 1)), dim=1).clone()     # events on step 2 and mid-way
    z_jump, v_jump = torch.randn(B, 2, Z) * 0.2, torch.randn(B, 2, V) * 0.2
    wx, wi = O.integrate_dae(method, _params(de.x_dot), _params(ae.i_calculator), x_init, t, x, z, v, i, a0, event_t, z_jump, v_jump)
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[method]
    ev = DAE_Event()
    ev.set_event(t=event_t.to(dev), z=z_jump.to(dev), v=v_jump.to(dev))
    with torch.no_grad():
        gx, gi = S().integrate_DAE(x_init=x_init.to(dev), x_func=de.to(dev), i_func=ae.to(dev), t=t.to(dev), x=x.to(dev), z=z.to(dev),
                                   v=v.to(dev), i=i.to(dev), all_initial=a0.to(dev), event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
    assert _native.last_kernel().startswith("psn_tc8_dae_kernel")
    assert torch.allclose(gx.cpu(), wx, rtol=RTOL, atol=ATOL), "x: " + tol_report(gx.cpu(), wx)
    assert torch.allclose(gi.cpu(), wi, rtol=RTOL, atol=ATOL), "i: " + tol_report(gi.cpu(), wi)


@pytest.mark.parametrize("X", [16, 7])
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_dae_tc_tape_gradients_all_methods(native_lib, method, X, monkeypatch):
    """DAE reverse sweep on tensor cores (ragged batch, an event on the very first step and one mid-way) against float64
    autograd through the oracle: both nets' parameters, x_init and all_initial."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, Euler, Midpoint, RK4, _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    torch.manual_seed(53)
    dev = "cuda:0"
    B, N, Z, V, I, H = 40, 24, 1, 2, 4, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z)
    t = (torch.arange(T, dtype=torch.float32) * 0.02).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda wd: torch.randn(T, B, wd) * 0.2
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    x_init = torch.randn(B, X) * 0.2
    a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
    event_t = torch.cat((t[0].view(B, 1, 1), t[N // 2].view(B, 1, 1)), dim=1).clone()     # events on step 1 and mid-way
    z_jump, v_jump = torch.randn(B, 2, Z) * 0.2, torch.randn(B, 2, V) * 0.2
    wx, wi = torch.randn(T, B, X) * 0.1, torch.randn(T, B, I) * 0.1
    pd = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    pa = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(ae.i_calculator)]
    xi64, a064 = x_init.double().requires_grad_(True), a0.double().requires_grad_(True)
    sx, si = O.integrate_dae(method, pd, pa, xi64, t.double(), x.double(), z.double(), v.double(), i.double(), a064,
                             event_t.double(), z_jump.double(), v_jump.double())
    ((sx * wx.double()).sum() + (si * wi.double()).sum()).backward()
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[method]
    de_d, ae_d = de.to(dev), ae.to(dev)
    ev = DAE_Event()
    ev.set_event(t=event_t.to(dev), z=z_jump.to(dev), v=v_jump.to(dev))
    xid, a0d = x_init.to(dev).requires_grad_(True), a0.to(dev).requires_grad_(True)
    gx, gi = S().integrate_DAE(x_init=xid, x_func=de_d, i_func=ae_d, t=t.to(dev), x=x.to(dev), z=z.to(dev), v=v.to(dev), i=i.to(dev),
                               all_initial=a0d, event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
    assert _native.last_kernel().startswith("psn_tc8_dae_kernel") and "tape" in _native.last_kernel()
    assert torch.allclose(gx.detach().cpu(), sx.detach().float(), rtol=RTOL, atol=ATOL)
    ((gx * wx.to(dev)).sum() + (gi * wi.to(dev)).sum()).backward()
    assert _native.last_kernel() == "psn_tc_dae_grad_reduce_kernel"
    lin_d = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    lin_a = [m for m in ae_d.i_calculator if isinstance(m, torch.nn.Linear)]
    pairs = [(lin_d[k].weight.grad, pd[k][0].grad) for k in range(4)] + [(lin_d[k].bias.grad, pd[k][1].grad) for k in range(4)]
    pairs += [(lin_a[k].weight.grad, pa[k][0].grad) for k in range(4)] + [(lin_a[k].bias.grad, pa[k][1].grad) for k in range(4)]
    pairs += [(xid.grad, xi64.grad), (a0d.grad, a064.grad)]
    for k, (g, g64) in enumerate(pairs):
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 1e-5 * scale + 1e-7, f"{method} tensor {k}: err {err:.3e} scale {scale:.3e}"


@pytest.mark.parametrize("T", [1, 2, 3])
@pytest.mark.parametrize("kind", ["ode", "dae"])
def test_tc_degenerate_grids(native_lib, kind, T, monkeypatch):
    """Grids with 0, 1 and 2 steps (the reference's loop simply does not iterate for T = 1): tensor-core forward + tape sweeps
    against float64 autograd through the oracle."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DE_Func, RK4
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    torch.manual_seed(70 + T)
    dev = "cuda:0"
    B, X, Z, V, I, H = 20, 16, 1, 2, 4, 64
    dae = kind == "dae"
    t = (torch.arange(T, dtype=torch.float32) * 0.05).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda wd: torch.randn(T, B, wd) * 0.2
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    wx, wi = torch.randn(T, B, X) * 0.1, torch.randn(T, B, I) * 0.1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V if dae else 0, i_dim=I if dae else 0)
    pd = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    if dae:
        ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z)
        pa = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(ae.i_calculator)]
        x_init = torch.randn(B, X) * 0.2
        a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
        xi64, a064 = x_init.double().requires_grad_(True), a0.double().requires_grad_(True)
        sx, si = O.integrate_dae("rk4", pd, pa, xi64, t.double(), x.double(), z.double(), v.double(), i.double(), a064)
        ((sx * wx.double()).sum() + (si * wi.double()).sum()).backward()
        de_d, ae_d = de.to(dev), ae.to(dev)
        xid, a0d = x_init.to(dev).requires_grad_(True), a0.to(dev).requires_grad_(True)
        gx, gi = RK4().integrate_DAE(x_init=xid, x_func=de_d, i_func=ae_d, t=t.to(dev), x=x.to(dev), z=z.to(dev), v=v.to(dev),
                                     i=i.to(dev), all_initial=a0d)
        assert torch.allclose(gx.detach().cpu(), sx.detach().float(), rtol=RTOL, atol=ATOL)
        assert torch.allclose(gi.detach().cpu(), si.detach().float(), rtol=RTOL, atol=ATOL)
        ((gx * wx.to(dev)).sum() + (gi * wi.to(dev)).sum()).backward()
        lin_a = [m for m in ae_d.i_calculator if isinstance(m, torch.nn.Linear)]
        pairs = [(lin_a[k].weight.grad, pa[k][0].grad) for k in range(4)] + [(lin_a[k].bias.grad, pa[k][1].grad) for k in range(4)]
        pairs += [(xid.grad, xi64.grad), (a0d.grad, a064.grad)]
    else:
        a0 = torch.cat((x[0], z[0]), dim=-1)
        x64, a064 = x.double().requires_grad_(True), a0.double().requires_grad_(True)
        sx = O.integrate_ode("rk4", pd, t.double(), x64, z.double(), a064)
        (sx * wx.double()).sum().backward()
        de_d = de.to(dev)
        xd, a0d = x.to(dev).requires_grad_(True), a0.to(dev).requires_grad_(True)
        gx = RK4().integrate_ODE(x_func=de_d, t=t.to(dev), x=xd, z=z.to(dev), all_initial=a0d)
        assert torch.allclose(gx.detach().cpu(), sx.detach().float(), rtol=RTOL, atol=ATOL)
        (gx * wx.to(dev)).sum().backward()
        pairs = [(xd.grad, x64.grad), (a0d.grad if a0d.grad is not None else torch.zeros_like(a0d), a064.grad if a064.grad is not None else torch.zeros_like(a064))]
    lin_d = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    for k in range(4):
        gW = lin_d[k].weight.grad if lin_d[k].weight.grad is not None else torch.zeros_like(lin_d[k].weight)
        gW64 = pd[k][0].grad if pd[k][0].grad is not None else torch.zeros_like(pd[k][0])
        pairs.append((gW, gW64))
    for k, (g, g64) in enumerate(pairs):
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 1e-5 * scale + 1e-7, f"{kind} T={T} tensor {k}: err {err:.3e} scale {scale:.3e}"


def test_four_warp_ab_kernel_refuses_narrow_state(native_lib):
    """Only the 8-warp kernel carries the X < 16 zero-padding; the 4-warp A/B build (impl="tc") must refuse, not mis-integrate."""
    from py_psnode_b200 import DE_Func, RK4, _native
    torch.manual_seed(3)
    dev = "cuda:0"
    B, T, X, Z = 16, 5, 6, 2
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=64).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.02).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, X, device=dev) * 0.2
    z = torch.randn(T, B, Z, device=dev) * 0.2
    a0 = torch.cat((x[0], z[0]), dim=-1)
    with torch.no_grad():
        RK4(impl="tc8").integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0)
        assert _native.last_kernel().startswith("psn_tc8_ode_kernel")
        with pytest.raises(RuntimeError):
            RK4(impl="tc").integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0)
