"""Parity at BASELINE.json's FULL sizes (configs[1] and configs[2]: B = 4096 trajectories x 1000 RK4 steps).

The CPU oracle cannot integrate 4096 x 1000 in seconds, but trajectories are independent units of the path (every op is
row-wise over the batch; SURVEY 8c item 5: the reference's outputs are batch-slice invariant), so the full-size GPU result
is checked row by row on a random SAMPLE of trajectories that the oracle integrates over all 1000 steps at
rtol=1e-5 / atol=1e-6, plus size-independent properties: x_sol[0] equals the given initial state bit for bit, two independent
kernel families agree, and the tape-based and the recomputing reverse sweeps give the same gradients."""
import numpy as np
import pytest
import torch

from helpers import ATOL, RTOL, tol_report

pytestmark = pytest.mark.gpu

B, N, X, H = 4096, 1000, 16, 64
T = N + 1


def _inputs(seed, widths, dev):
    g = torch.Generator().manual_seed(seed)
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    out = {"t": t}
    for name, w in widths.items():
        out[name] = torch.randn(T, B, w, generator=g) * 0.1
    out["x0"] = torch.randn(B, X, generator=g) * 0.1
    return out, {k: v.to(dev) for k, v in out.items()}


def _params(mod):
    return [(m.weight.detach().cpu(), m.bias.detach().cpu()) for m in mod if isinstance(m, torch.nn.Linear)]


def test_cfg2_forward_full_size_against_oracle_sample(native_lib):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, ODE_Event, RK4, _native
    dev = "cuda:0"
    torch.manual_seed(11)
    Z = 2
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H)
    host, d = _inputs(12, {"z": Z}, dev)
    event_t = host["t"][N // 2].view(B, 1, 1).clone()
    z_jump = torch.randn(B, 1, Z) * 0.1
    ev = ODE_Event()
    ev.set_event(t=event_t.to(dev), z=z_jump.to(dev))
    x_view = d["x0"].unsqueeze(0).expand(T, B, X)
    a0 = torch.cat((d["x0"], d["z"][0]), dim=-1)
    de_d = de.to(dev)
    outs = {}
    with torch.no_grad():
        for impl in ("auto", "fused"):
            outs[impl] = RK4(impl=impl).integrate_ODE(x_func=de_d, t=d["t"], x=x_view, z=d["z"], all_initial=a0,
                                                      event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
            if impl == "auto":
                assert _native.last_kernel().startswith("psn_tc8_ode_kernel")
    got = outs["auto"].cpu()
    assert torch.equal(got[0], host["x0"]), "x_sol[0] must be the initial state, bit for bit"
    assert torch.isfinite(got).all()
    # independent kernel families (tcgen05 3xTF32 vs CUDA-core fp32 FMA) over the full batch
    assert torch.allclose(outs["auto"], outs["fused"], rtol=3 * RTOL, atol=3 * ATOL), tol_report(got, outs["fused"].cpu())
    # the oracle on a sample of trajectories, all 1000 steps (event at step 500 included)
    rows = torch.from_numpy(np.random.default_rng(0).choice(B, size=48, replace=False)).sort().values
    rows[0], rows[-1] = 0, B - 1
    de_c = de.cpu()
    xs = host["x0"][rows].unsqueeze(0).expand(T, len(rows), X)
    a0s = torch.cat((host["x0"][rows], host["z"][0][rows]), dim=-1)
    want = O.integrate_ode("rk4", _params(de_c.x_dot), host["t"][:, rows], xs, host["z"][:, rows], a0s, event_t[rows], z_jump[rows])
    assert torch.allclose(got[:, rows], want, rtol=RTOL, atol=ATOL), tol_report(got[:, rows], want)


def test_cfg3_dae_forward_full_size_against_oracle_sample(native_lib):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4, _native
    dev = "cuda:0"
    torch.manual_seed(21)
    Z, V, I = 1, 2, 4
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z)
    host, d = _inputs(22, {"z": Z, "v": V, "i": I}, dev)
    event_t = host["t"][N // 3].view(B, 1, 1).clone()
    z_jump, v_jump = torch.randn(B, 1, Z) * 0.1, torch.randn(B, 1, V) * 0.1
    ev = DAE_Event()
    ev.set_event(t=event_t.to(dev), z=z_jump.to(dev), v=v_jump.to(dev))
    x_view = d["x0"].unsqueeze(0).expand(T, B, X)
    a0 = torch.cat((d["x0"], d["z"][0], d["v"][0], d["i"][0]), dim=-1)
    with torch.no_grad():
        gx, gi = RK4().integrate_DAE(x_init=d["x0"], x_func=de.to(dev), i_func=ae.to(dev), t=d["t"], x=x_view, z=d["z"], v=d["v"],
                                     i=d["i"], all_initial=a0, event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
    assert _native.last_kernel().startswith("psn_tc8_dae_kernel")
    gx, gi = gx.cpu(), gi.cpu()
    assert torch.equal(gx[0], host["x0"])
    assert torch.isfinite(gx).all() and torch.isfinite(gi).all()
    rows = torch.from_numpy(np.random.default_rng(1).choice(B, size=32, replace=False)).sort().values
    rows[0], rows[-1] = 0, B - 1
    de_c, ae_c = de.cpu(), ae.cpu()
    xs = host["x0"][rows].unsqueeze(0).expand(T, len(rows), X)
    a0s = torch.cat((host["x0"][rows], host["z"][0][rows], host["v"][0][rows], host["i"][0][rows]), dim=-1)
    wx, wi = O.integrate_dae("rk4", _params(de_c.x_dot), _params(ae_c.i_calculator), host["x0"][rows], host["t"][:, rows], xs,
                             host["z"][:, rows], host["v"][:, rows], host["i"][:, rows], a0s, event_t[rows], z_jump[rows], v_jump[rows])
    assert torch.allclose(gx[:, rows], wx, rtol=RTOL, atol=ATOL), "x: " + tol_report(gx[:, rows], wx)
    assert torch.allclose(gi[:, rows], wi, rtol=RTOL, atol=ATOL), "i: " + tol_report(gi[:, rows], wi)


def test_cfg2_gradients_full_size_tape_vs_recompute(native_lib, monkeypatch):
    """4.1 M traj-steps: the tensor-core reverse sweep (activation tape) and the generic recomputing sweep are two independent
    implementations of the same discrete adjoint; their parameter / initial-state gradients must agree to fp32 summation noise."""
    from py_psnode_b200 import DE_Func, RK4, _native
    dev = "cuda:0"
    torch.manual_seed(31)
    Z = 2
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H).to(dev)
    _, d = _inputs(32, {"z": Z}, dev)
    w = torch.randn(T, B, X, device=dev) * 1e-3                     # dL/dx_sol
    grads = {}
    for sweep in ("tape", "recompute"):
        if sweep == "recompute":
            monkeypatch.setenv("PSNODE_TAPE_MAX_GB", "0")
        else:
            monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
        for p in de.parameters():
            p.grad = None
        x0 = d["x0"].clone().requires_grad_(True)
        x_view = x0.unsqueeze(0).expand(T, B, X)
        a0 = torch.cat((x0.detach(), d["z"][0]), dim=-1).requires_grad_(True)
        sol = RK4().integrate_ODE(x_func=de, t=d["t"], x=x_view, z=d["z"], all_initial=a0)
        sol.backward(w)
        assert _native.last_kernel() == ("psn_tc_grad_reduce_kernel" if sweep == "tape" else "psn_grad_reduce_kernel")
        grads[sweep] = [p.grad.clone() for p in de.parameters()] + [x0.grad.clone(), a0.grad.clone()]
    for k, (a, b) in enumerate(zip(grads["tape"], grads["recompute"])):
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        assert err <= 2e-5 * scale + 1e-9, f"tensor {k}: max|tape - recompute| = {err:.3e}, scale {scale:.3e}"


def test_cfg3_gradients_full_size_tape_vs_recompute(native_lib, monkeypatch):
    """cfg3 (DAE) at full size: tensor-core reverse sweep (DAE tape) against the generic recomputing sweep."""
    from py_psnode_b200 import AE_Func, DE_Func, RK4, _native
    dev = "cuda:0"
    torch.manual_seed(33)
    Z, V, I = 1, 2, 4
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I).to(dev)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z).to(dev)
    _, d = _inputs(34, {"z": Z, "v": V, "i": I}, dev)
    wx = torch.randn(T, B, X, device=dev) * 1e-3
    wi = torch.randn(T, B, I, device=dev) * 1e-3
    plist = list(de.parameters()) + list(ae.parameters())
    grads = {}
    for sweep in ("tape", "recompute"):
        if sweep == "recompute":
            monkeypatch.setenv("PSNODE_TAPE_MAX_GB", "0")
        else:
            monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
        for p in plist:
            p.grad = None
        x0 = d["x0"].clone().requires_grad_(True)
        a0 = torch.cat((x0.detach(), d["z"][0], d["v"][0], d["i"][0]), dim=-1).requires_grad_(True)
        x_view = d["x0"].unsqueeze(0).expand(T, B, X)
        gx, gi = RK4().integrate_DAE(x_init=x0, x_func=de, i_func=ae, t=d["t"], x=x_view, z=d["z"], v=d["v"], i=d["i"], all_initial=a0)
        ((gx * wx).sum() + (gi * wi).sum()).backward()
        assert _native.last_kernel() == ("psn_tc_dae_grad_reduce_kernel" if sweep == "tape" else "psn_grad_reduce_kernel")
        grads[sweep] = [p.grad.clone() for p in plist] + [x0.grad.clone(), a0.grad.clone()]
    for k, (a, b) in enumerate(zip(grads["tape"], grads["recompute"])):
        scale = float(b.abs().max())
        err = float((a - b).abs().max())
        assert err <= 2e-5 * scale + 1e-9, f"tensor {k}: max|tape - recompute| = {err:.3e}, scale {scale:.3e}"
