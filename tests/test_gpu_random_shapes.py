"""Seeded random sweep over the shapes the tensor-core kernels accept (X <= 16 states, up to 8 held-input columns, ragged
batches, 0..2 events, non-uniform per-trajectory grids, all three schemes): forward against the fp32 oracle at
rtol=1e-5 / atol=1e-6, tape-based reverse sweep against float64 autograd through the oracle."""
import random

import pytest
import torch

from helpers import ATOL, RTOL, tol_report

pytestmark = pytest.mark.gpu


def _params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]


def _grid(T, B, rng):
    """per-trajectory step sizes (the reference takes dt = t[j] - t[j-1] per sample), shared event instants"""
    base = torch.tensor([rng.uniform(0.005, 0.03) for _ in range(T - 1)], dtype=torch.float32)
    t0 = torch.cat((torch.zeros(1), torch.cumsum(base, 0)))
    t = t0.view(T, 1, 1).repeat(1, B, 1).clone()
    t[:, 1:, 0] += torch.linspace(0.0, 1e-3, T).view(T, 1) * torch.rand(1, B - 1) if B > 1 else 0.0   # sample 0 keeps the event times exact
    return t


@pytest.mark.parametrize("seed", range(8))
def test_ode_random_shapes(native_lib, seed, monkeypatch):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, Euler, Midpoint, ODE_Event, RK4, _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    rng = random.Random(1000 + seed)
    torch.manual_seed(2000 + seed)
    dev = "cuda:0"
    X, Z = rng.randint(1, 16), rng.randint(1, 8)
    B, T = rng.randint(1, 70), rng.randint(2, 26)
    E = rng.randint(0, 2) if T > 3 else 0
    method = rng.choice(["euler", "midpoint", "rk4"])
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=64)
    t = _grid(T, B, rng)
    x, z = torch.randn(T, B, X) * 0.2, torch.randn(T, B, Z) * 0.2
    w = torch.randn(T, B, X) * 0.1
    a0 = torch.cat((x[0], z[0]), dim=-1)
    ev_args, ev = (None, None), None
    if E:
        steps = sorted(rng.sample(range(0, T - 1), E))
        event_t = torch.stack([t[s] for s in steps], dim=1).clone()           # (B, E, 1): sample 0's times are the ones tested
        z_jump = torch.randn(B, E, Z) * 0.2
        ev_args = (event_t, z_jump)
    want = O.integrate_ode(method, _params(de.x_dot), t, x, z, a0, *ev_args)
    p64 = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    a064, x64 = a0.double().requires_grad_(True), x.double().requires_grad_(True)
    s64 = O.integrate_ode(method, p64, t.double(), x64, z.double(), a064, *(a.double() if a is not None else None for a in ev_args))
    (s64 * w.double()).sum().backward()
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[method]
    de_d = de.to(dev)
    kw = {}
    if E:
        ev = ODE_Event()
        ev.set_event(t=ev_args[0].to(dev), z=ev_args[1].to(dev))
        kw = dict(event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
    xd, a0d = x.to(dev).requires_grad_(True), a0.to(dev).requires_grad_(True)
    sol = S().integrate_ODE(x_func=de_d, t=t.to(dev), x=xd, z=z.to(dev), all_initial=a0d, **kw)
    desc = f"X={X} Z={Z} B={B} T={T} E={E} {method}"
    assert _native.last_kernel().startswith("psn_tc8_ode_kernel"), desc
    assert torch.allclose(sol.detach().cpu(), want, rtol=RTOL, atol=ATOL), desc + " " + tol_report(sol.detach().cpu(), want)
    (sol * w.to(dev)).sum().backward()
    assert _native.last_kernel() == "psn_tc_grad_reduce_kernel", desc
    lin = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    pairs = [(lin[k].weight.grad, p64[k][0].grad) for k in range(4)] + [(lin[k].bias.grad, p64[k][1].grad) for k in range(4)]
    pairs += [(a0d.grad, a064.grad), (xd.grad[0], x64.grad[0])]
    for k, (g, g64) in enumerate(pairs):
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 1e-5 * scale + 1e-7, f"{desc} tensor {k}: err {err:.3e} scale {scale:.3e}"


@pytest.mark.parametrize("seed", range(6))
def test_dae_random_shapes(native_lib, seed, monkeypatch):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, Euler, Midpoint, RK4, _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    rng = random.Random(3000 + seed)
    torch.manual_seed(4000 + seed)
    dev = "cuda:0"
    X = rng.randint(1, 16)
    Z, V = rng.randint(1, 3), rng.randint(1, 3)
    I = rng.randint(1, 8 - Z - V)
    B, T = rng.randint(1, 60), rng.randint(2, 22)
    E = rng.randint(0, 2) if T > 3 else 0
    method = rng.choice(["euler", "midpoint", "rk4"])
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=64, v_dim=V, i_dim=I)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=64, z_dim=Z)
    t = _grid(T, B, rng)
    mk = lambda wd: torch.randn(T, B, wd) * 0.2
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    x_init = torch.randn(B, X) * 0.2
    a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
    ev_args = (None, None, None)
    if E:
        steps = sorted(rng.sample(range(0, T - 1), E))
        ev_args = (torch.stack([t[s] for s in steps], dim=1).clone(), torch.randn(B, E, Z) * 0.2, torch.randn(B, E, V) * 0.2)
    wx, wi = torch.randn(T, B, X) * 0.1, torch.randn(T, B, I) * 0.1
    want_x, want_i = O.integrate_dae(method, _params(de.x_dot), _params(ae.i_calculator), x_init, t, x, z, v, i, a0, *ev_args)
    pd = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    pa = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(ae.i_calculator)]
    xi64, a064 = x_init.double().requires_grad_(True), a0.double().requires_grad_(True)
    sx, si = O.integrate_dae(method, pd, pa, xi64, t.double(), x.double(), z.double(), v.double(), i.double(), a064,
                             *(a.double() if a is not None else None for a in ev_args))
    ((sx * wx.double()).sum() + (si * wi.double()).sum()).backward()
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[method]
    de_d, ae_d = de.to(dev), ae.to(dev)
    kw = {}
    if E:
        ev = DAE_Event()
        ev.set_event(t=ev_args[0].to(dev), z=ev_args[1].to(dev), v=ev_args[2].to(dev))
        kw = dict(event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
    xid, a0d = x_init.to(dev).requires_grad_(True), a0.to(dev).requires_grad_(True)
    gx, gi = S().integrate_DAE(x_init=xid, x_func=de_d, i_func=ae_d, t=t.to(dev), x=x.to(dev), z=z.to(dev), v=v.to(dev), i=i.to(dev),
                               all_initial=a0d, **kw)
    desc = f"X={X} Z={Z} V={V} I={I} B={B} T={T} E={E} {method}"
    assert _native.last_kernel().startswith("psn_tc8_dae_kernel"), desc
    assert torch.allclose(gx.detach().cpu(), want_x, rtol=RTOL, atol=ATOL), desc + " x: " + tol_report(gx.detach().cpu(), want_x)
    assert torch.allclose(gi.detach().cpu(), want_i, rtol=RTOL, atol=ATOL), desc + " i: " + tol_report(gi.detach().cpu(), want_i)
    ((gx * wx.to(dev)).sum() + (gi * wi.to(dev)).sum()).backward()
    assert _native.last_kernel() == "psn_tc_dae_grad_reduce_kernel", desc
    lin_d = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    lin_a = [m for m in ae_d.i_calculator if isinstance(m, torch.nn.Linear)]
    pairs = [(lin_d[k].weight.grad, pd[k][0].grad) for k in range(4)] + [(lin_d[k].bias.grad, pd[k][1].grad) for k in range(4)]
    pairs += [(lin_a[k].weight.grad, pa[k][0].grad) for k in range(4)] + [(lin_a[k].bias.grad, pa[k][1].grad) for k in range(4)]
    pairs += [(xid.grad, xi64.grad), (a0d.grad, a064.grad)]
    for k, (g, g64) in enumerate(pairs):
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 1e-5 * scale + 1e-7, f"{desc} tensor {k}: err {err:.3e} scale {scale:.3e}"
