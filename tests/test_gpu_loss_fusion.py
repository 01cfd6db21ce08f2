"""Integration fused with the scripts' masked squared-error loss (neural_00_ODE_01_no_encode.py:353-355,
neural_01_DAE_01_no_encode.py:414-418): `integrate_ODE_loss` / `integrate_DAE_loss` must give the same loss value and the same
gradients as integrate_* followed by the loss, on every sweep family -- the tensor-core sweeps form dL/dx_sol inside the
sweep (psnode_adjoint.fuse_x / fuse_i), the recomputing sweep gets it from the fused loss kernel."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _grads(params, extra):
    return [p.grad.clone() for p in params] + [e.grad.clone() for e in extra]


def _close(name, a, b):
    scale = float(b.abs().max())
    err = float((a - b).abs().max())
    assert err <= 2e-5 * scale + 1e-9, f"{name}: max|fused - unfused| = {err:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("H,X,Z,depth,kernel", [(64, 16, 2, 4, "psn_tc_grad_reduce_kernel"), (128, 128, 128, 2, "psn_wide_assemble_kernel"),
                                                (32, 5, 2, 4, "psn_grad_reduce_kernel")])
def test_ode_loss_fusion_matches_unfused(native_lib, H, X, Z, depth, kernel):
    from py_psnode_b200 import DE_Func, RK4, _native
    from py_psnode_b200.losses import masked_sse
    torch.manual_seed(81)
    B, N = 40, 18
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, depth=depth).to(DEV)
    t = (torch.arange(T, dtype=torch.float32, device=DEV) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    z = torch.randn(T, B, Z, device=DEV) * 0.1
    # batch-major storage viewed time-major, as the scripts pass it
    target = (torch.randn(B, T, X, device=DEV) * 0.1).permute(1, 0, 2)
    mask = torch.ones(B, T, 1, device=DEV)
    mask[:, T - 3:] = 0.0
    mask = mask.permute(1, 0, 2)
    fw = torch.ones(X, device=DEV)
    fw[1] = 10.0
    plist = list(de.parameters())
    res = {}
    for mode in ("unfused", "fused"):
        for p in plist:
            p.grad = None
        x0 = (target[0].clone() * 0.5).requires_grad_(True)
        a0 = torch.cat((x0.detach(), z[0]), dim=-1).requires_grad_(True)
        xv = x0.unsqueeze(0).expand(T, B, X)
        if mode == "unfused":
            sol = RK4().integrate_ODE(x_func=de, t=t, x=xv, z=z, all_initial=a0)
            num = masked_sse(sol, target, mask, fw)
        else:
            num, sol = RK4().integrate_ODE_loss(x_func=de, t=t, x=xv, z=z, all_initial=a0, target=target, mask=mask, feat_weight=fw)
            assert not sol.requires_grad
        (num / mask.sum() * 3.0).backward()
        assert _native.last_kernel() == kernel, _native.last_kernel()
        res[mode] = (num.detach().clone(), _grads(plist, [x0, a0]))
    assert torch.allclose(res["fused"][0], res["unfused"][0], rtol=1e-6)
    for k, (a, b) in enumerate(zip(res["fused"][1], res["unfused"][1])):
        _close(f"tensor {k}", a, b)


def test_dae_loss_fusion_matches_unfused(native_lib):
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4, _native
    from py_psnode_b200.losses import masked_sse
    torch.manual_seed(82)
    B, N, X, Z, V, I, H = 36, 15, 16, 1, 2, 4, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I).to(DEV)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z).to(DEV)
    t = (torch.arange(T, dtype=torch.float32, device=DEV) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda w: torch.randn(T, B, w, device=DEV) * 0.1
    z, v, i, tx, ti = mk(Z), mk(V), mk(I), mk(X), mk(I)
    mask = torch.ones(T, B, 1, device=DEV)
    mask[T - 2:] = 0.0
    fwx = torch.ones(X, device=DEV)
    fwx[1] = 10.0
    ev = DAE_Event()
    ev.set_event(t=t[N // 2].view(B, 1, 1).clone(), z=torch.randn(B, 1, Z, device=DEV) * 0.1, v=torch.randn(B, 1, V, device=DEV) * 0.1)
    plist = list(de.parameters()) + list(ae.parameters())
    res = {}
    for mode in ("unfused", "fused"):
        for p in plist:
            p.grad = None
        x0 = (tx[0].clone() * 0.5).requires_grad_(True)
        a0 = torch.cat((x0.detach(), z[0], v[0], i[0]), dim=-1).requires_grad_(True)
        kw = dict(x_init=x0, x_func=de, i_func=ae, t=t, x=tx, z=z, v=v, i=i, all_initial=a0, event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
        if mode == "unfused":
            sx, si = RK4().integrate_DAE(**kw)
            num = masked_sse(sx, tx, mask, fwx) + masked_sse(si, ti, mask)
        else:
            num, sx, si = RK4().integrate_DAE_loss(target_x=tx, target_i=ti, mask=mask, feat_weight_x=fwx, **kw)
        (num / mask.sum()).backward()
        assert _native.last_kernel() == "psn_tc_dae_grad_reduce_kernel", _native.last_kernel()
        res[mode] = (num.detach().clone(), _grads(plist, [x0, a0]))
    assert torch.allclose(res["fused"][0], res["unfused"][0], rtol=1e-6)
    for k, (a, b) in enumerate(zip(res["fused"][1], res["unfused"][1])):
        _close(f"tensor {k}", a, b)


def test_layer_dae_loss_fusion_matches_unfused(native_lib):
    """Latent DAE_02 net on the layer path (H = 128): the recomputing sweep forms the upstream rows of both loss terms on the fly."""
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4, _native
    from py_psnode_b200.losses import masked_sse
    torch.manual_seed(83)
    B, N, H = 36, 13, 128
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(DEV)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(DEV)
    t = (torch.arange(T, dtype=torch.float32, device=DEV) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: torch.randn(T, B, H, device=DEV) * 0.05
    z0, v0, i, tx, ti = mk(), mk(), mk(), mk(), mk()
    mask = torch.ones(T, B, 1, device=DEV)
    mask[T - 2:] = 0.0
    fwx = torch.ones(H, device=DEV)
    fwx[1] = 10.0
    ev = DAE_Event()
    ev.set_event(t=t[N // 2].view(B, 1, 1).clone(), z=torch.randn(B, 1, H, device=DEV) * 0.05, v=torch.randn(B, 1, H, device=DEV) * 0.05)
    plist = list(de.parameters()) + list(ae.parameters())
    res = {}
    for mode in ("unfused", "fused"):
        for p in plist:
            p.grad = None
        z, v = z0.clone().requires_grad_(True), v0.clone().requires_grad_(True)
        x0 = (tx[0].clone() * 0.5).requires_grad_(True)
        a0 = torch.cat((x0.detach(), z0[0], v0[0], i[0]), dim=-1).requires_grad_(True)
        kw = dict(x_init=x0, x_func=de, i_func=ae, t=t, x=tx, z=z, v=v, i=i, all_initial=a0, event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
        if mode == "unfused":
            sx, si = RK4().integrate_DAE(**kw)
            num = masked_sse(sx, tx, mask, fwx) + masked_sse(si, ti, mask)
        else:
            num, sx, si = RK4().integrate_DAE_loss(target_x=tx, target_i=ti, mask=mask, feat_weight_x=fwx, **kw)
        (num / mask.sum()).backward()
        assert _native.last_kernel().startswith("psn_lg_"), _native.last_kernel()
        res[mode] = (num.detach().clone(), _grads(plist, [x0, a0, z, v]))
    assert torch.allclose(res["fused"][0], res["unfused"][0], rtol=1e-6)
    for k, (a, b) in enumerate(zip(res["fused"][1], res["unfused"][1])):
        _close(f"tensor {k}", a, b)
