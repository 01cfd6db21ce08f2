"""BASELINE configs[3] and configs[4] shapes at reduced batch / length: the `*_02_direct_encode` models integrate in a
latent space whose width equals the hidden width (neural_00_ODE_02_direct_encode.py:70: X = Z = H = 128;
neural_01_DAE_02_direct_encode.py:61-121: X = Z = V = I = H = 256) with 2-layer nets whose first layer (3S x H, up to
3.1 MB) does not fit one SM's shared memory.  Forward against the CPU oracle at rtol=1e-5 / atol=1e-6, gradients (with the
input-series gradients the encoders need) against torch autograd through the oracle."""
import pytest
import torch

from helpers import ATOL, RTOL, tol_report

pytestmark = pytest.mark.gpu


def _params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]


@pytest.mark.parametrize("solver", ["euler", "rk4"])
def test_cfg4_shape_latent_ode_h128(native_lib, solver):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, Euler, RK4, _native
    torch.manual_seed(41)
    dev = "cuda:0"
    B, N, H = 48, 24, 128
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, H) * 0.1
    z = torch.randn(T, B, H) * 0.1
    a0 = torch.cat((x[0], z[0]), dim=-1)
    want = O.integrate_ode(solver, _params(de.x_dot), t, x, z, a0)
    S = {"euler": Euler, "rk4": RK4}[solver]
    with torch.no_grad():
        got = S().integrate_ODE(x_func=de.to(dev), t=t.to(dev), x=x.to(dev), z=z.to(dev), all_initial=a0.to(dev)).cpu()
    assert _native.last_kernel().startswith("psn_wide_fwd_kernel")      # round 2: tcgen05 kernels for the latent widths
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want)


def test_cfg4_shape_gradients_with_latent_input_grads(native_lib):
    """Adjoint training of the encoded ODE model: the latent inputs Zh (T,B,H) and x[0] carry gradients back to the encoders."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, RK4
    torch.manual_seed(42)
    dev = "cuda:0"
    B, N, H = 16, 12, 128
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, H) * 0.1
    z = torch.randn(T, B, H) * 0.1
    w = torch.randn(T, B, H) * 0.1
    # oracle (float64 autograd) ------------------------------------------------------------------
    p64 = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    x64, z64 = x.double().requires_grad_(True), z.double().requires_grad_(True)
    a064 = torch.cat((x64[0], z64[0]), dim=-1)
    sol64 = O.integrate_ode("rk4", p64, t.double(), x64, z64, a064)
    (sol64 * w.double()).sum().backward()
    # CUDA path -----------------------------------------------------------------------------------
    de_d = de.to(dev)
    xd, zd = x.to(dev).requires_grad_(True), z.to(dev).requires_grad_(True)
    a0d = torch.cat((xd[0], zd[0]), dim=-1)
    sol = RK4().integrate_ODE(x_func=de_d, t=t.to(dev), x=xd, z=zd, all_initial=a0d)
    (sol * w.to(dev)).sum().backward()
    lin = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    pairs = [(lin[k].weight.grad, p64[k][0].grad) for k in range(2)] + [(lin[k].bias.grad, p64[k][1].grad) for k in range(2)]
    pairs += [(xd.grad, x64.grad), (zd.grad, z64.grad)]
    for k, (g, g64) in enumerate(pairs):
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 2e-5 * scale + 1e-7, f"tensor {k}: err {err:.3e} scale {scale:.3e}"


def test_cfg5_shape_latent_dae_h256(native_lib):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4, _native
    torch.manual_seed(43)
    dev = "cuda:0"
    B, N, H = 24, 10, 256
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: torch.randn(T, B, H) * 0.05
    x, z, v, i = mk(), mk(), mk(), mk()
    x_init = torch.randn(B, H) * 0.05
    a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1)
    event_t = t[N // 2].view(B, 1, 1).clone()
    z_jump, v_jump = torch.randn(B, 1, H) * 0.05, torch.randn(B, 1, H) * 0.05
    wx, wi = O.integrate_dae("rk4", _params(de.x_dot), _params(ae.i_calculator), x_init, t, x, z, v, i, a0, event_t, z_jump, v_jump)
    ev = DAE_Event()
    ev.set_event(t=event_t.to(dev), z=z_jump.to(dev), v=v_jump.to(dev))
    with torch.no_grad():
        gx, gi = RK4().integrate_DAE(x_init=x_init.to(dev), x_func=de.to(dev), i_func=ae.to(dev), t=t.to(dev), x=x.to(dev), z=z.to(dev),
                                     v=v.to(dev), i=i.to(dev), all_initial=a0.to(dev), event_fn=ev.event_fn,
                                     jump_change_fn=ev.jump_change_fn)
    assert _native.last_kernel().startswith("psn_lg_gemm_kernel")      # round 2: per-layer tcgen05 GEMMs for H = 256
    assert torch.allclose(gx.cpu(), wx, rtol=RTOL, atol=ATOL), "x: " + tol_report(gx.cpu(), wx)
    assert torch.allclose(gi.cpu(), wi, rtol=RTOL, atol=ATOL), "i: " + tol_report(gi.cpu(), wi)


def test_cfg5_shape_gradients_h256(native_lib):
    """Reverse sweep at the cfg5 widths (2 trajectories per CTA build of the generic sweep): parameter, x_init, all_initial and
    latent-input (Zh, Vh) gradients against float64 autograd through the oracle."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4, _native
    torch.manual_seed(44)
    dev = "cuda:0"
    B, N, H = 6, 6, 256
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: torch.randn(T, B, H) * 0.05
    x, z, v, i = mk(), mk(), mk(), mk()
    x_init = torch.randn(B, H) * 0.05
    wx, wi = torch.randn(T, B, H) * 0.1, torch.randn(T, B, H) * 0.1
    # oracle, float64 autograd
    pd = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(de.x_dot)]
    pa = [(W.double().requires_grad_(True), b.double().requires_grad_(True)) for W, b in _params(ae.i_calculator)]
    xi64, z64, v64 = x_init.double().requires_grad_(True), z.double().requires_grad_(True), v.double().requires_grad_(True)
    a064 = torch.cat((xi64, z64[0], v64[0], i.double()[0]), dim=-1)
    sx, si = O.integrate_dae("rk4", pd, pa, xi64, t.double(), x.double(), z64, v64, i.double(), a064)
    ((sx * wx.double()).sum() + (si * wi.double()).sum()).backward()
    # CUDA path
    de_d, ae_d = de.to(dev), ae.to(dev)
    xid = x_init.to(dev).requires_grad_(True)
    zd, vd = z.to(dev).requires_grad_(True), v.to(dev).requires_grad_(True)
    a0d = torch.cat((xid, zd[0], vd[0], i.to(dev)[0]), dim=-1)
    # impl = generic: the CUDA-core sweep stays covered at these widths (auto takes the layer path's tensor-core sweep, test_gpu_layer.py)
    gx, gi = RK4(impl="generic").integrate_DAE(x_init=xid, x_func=de_d, i_func=ae_d, t=t.to(dev), x=x.to(dev), z=zd, v=vd, i=i.to(dev), all_initial=a0d)
    ((gx * wx.to(dev)).sum() + (gi * wi.to(dev)).sum()).backward()
    assert _native.last_kernel() == "psn_grad_reduce_kernel"
    lin_d = [m for m in de_d.x_dot if isinstance(m, torch.nn.Linear)]
    lin_a = [m for m in ae_d.i_calculator if isinstance(m, torch.nn.Linear)]
    pairs = [(lin_d[k].weight.grad, pd[k][0].grad) for k in range(2)] + [(lin_d[k].bias.grad, pd[k][1].grad) for k in range(2)]
    pairs += [(lin_a[k].weight.grad, pa[k][0].grad) for k in range(2)] + [(lin_a[k].bias.grad, pa[k][1].grad) for k in range(2)]
    pairs += [(xid.grad, xi64.grad), (zd.grad, z64.grad), (vd.grad, v64.grad)]
    for k, (g, g64) in enumerate(pairs):
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 2e-5 * scale + 1e-7, f"tensor {k}: err {err:.3e} scale {scale:.3e}"
