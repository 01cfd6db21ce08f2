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
    de_c = de_d          # _params() copies the weights to the host; the module itself stays on the GPU
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


# ---------------------------------------------------------------------------------------------------------------------------
# Full-size GRADIENTS against the oracle (float64 autograd), not only against the repo's own second sweep:
#   * d_x0 / d_a0 are per-trajectory: the rows of a random sample of the full-batch run are compared with float64 autograd
#     through the oracle on exactly those trajectories (trajectories are independent, so their rows do not depend on the rest);
#   * parameter gradients are sums over trajectories: the GPU sweep on the sample alone (all 1000 steps) is compared with the
#     oracle's float64 gradient of the same sample, and the full-batch gradient with the sum over four 1024-trajectory chunks.
def _rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), 1e-30)


def test_cfg2_gradients_full_size_against_oracle_fp64_sample(native_lib, monkeypatch):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, RK4, _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    dev = "cuda:0"
    torch.manual_seed(61)
    Z = 2
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H)
    host, d = _inputs(62, {"z": Z}, dev)
    w = torch.randn(T, B, X, generator=torch.Generator().manual_seed(63)) * 1e-3
    de_d = de.to(dev)
    plist = list(de_d.parameters())

    def gpu_grads(rows):
        for p in plist:
            p.grad = None
        x0 = d["x0"][rows].clone().requires_grad_(True)
        nb = x0.shape[0]
        a0 = torch.cat((x0.detach(), d["z"][0, rows]), dim=-1).requires_grad_(True)
        sol = RK4().integrate_ODE(x_func=de_d, t=d["t"][:, rows], x=x0.unsqueeze(0).expand(T, nb, X), z=d["z"][:, rows], all_initial=a0)
        sol.backward(w[:, rows].to(dev))
        assert _native.last_kernel() == "psn_tc_grad_reduce_kernel"
        return [p.grad.clone().cpu() for p in plist], x0.grad.cpu(), a0.grad.cpu()

    full_theta, full_dx0, full_da0 = gpu_grads(slice(0, B))
    rows = torch.randperm(B, generator=torch.Generator().manual_seed(64))[:32].sort().values
    sub_theta, sub_dx0, sub_da0 = gpu_grads(rows)
    # oracle, float64 autograd on the sample over all 1000 steps
    p64 = [(Wm.double().requires_grad_(True), bm.double().requires_grad_(True)) for Wm, bm in _params(de_d.x_dot)]
    x064 = host["x0"][rows].double().requires_grad_(True)
    a064 = torch.cat((x064.detach(), host["z"][0, rows].double()), dim=-1).requires_grad_(True)
    sol64 = O.integrate_ode("rk4", p64, host["t"][:, rows].double(), x064.unsqueeze(0).expand(T, len(rows), X), host["z"][:, rows].double(), a064)
    (sol64 * w[:, rows].double()).sum().backward()
    ref_theta = [q for pair in p64 for q in (pair[0].grad, pair[1].grad)]
    report = [f"theta[{k}] rel {_rel(g, r):.2e}" for k, (g, r) in enumerate(zip(sub_theta, ref_theta))]
    report += [f"d_x0 (full-batch rows) rel {_rel(full_dx0[rows], x064.grad):.2e}", f"d_a0 (full-batch rows) rel {_rel(full_da0[rows], a064.grad):.2e}"]
    print("cfg2 full-length gradients vs oracle fp64:\n  " + "\n  ".join(report))
    for g, r in zip(sub_theta, ref_theta):
        assert _rel(g, r) <= 2e-5
    assert _rel(full_dx0[rows], x064.grad) <= 2e-5 and _rel(full_da0[rows], a064.grad) <= 2e-5
    # batch additivity of the parameter gradients at full size
    acc = None
    for c in range(4):
        th, _, _ = gpu_grads(slice(1024 * c, 1024 * (c + 1)))
        acc = th if acc is None else [a + b for a, b in zip(acc, th)]
    for k, (a, b) in enumerate(zip(acc, full_theta)):
        assert _rel(a, b) <= 2e-5, f"theta[{k}]: sum of chunk gradients vs full batch rel {_rel(a, b):.2e}"


def test_cfg3_gradients_full_size_against_oracle_fp64_sample(native_lib, monkeypatch):
    from oracle import psnode_oracle as O
    from py_psnode_b200 import AE_Func, DE_Func, RK4, _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    dev = "cuda:0"
    torch.manual_seed(65)
    Z, V, I = 1, 2, 4
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z)
    host, d = _inputs(66, {"z": Z, "v": V, "i": I}, dev)
    g = torch.Generator().manual_seed(67)
    wx, wi = torch.randn(T, B, X, generator=g) * 1e-3, torch.randn(T, B, I, generator=g) * 1e-3
    de_d, ae_d = de.to(dev), ae.to(dev)
    plist = list(de_d.parameters()) + list(ae_d.parameters())

    def gpu_grads(rows):
        for p in plist:
            p.grad = None
        x0 = d["x0"][rows].clone().requires_grad_(True)
        nb = x0.shape[0]
        a0 = torch.cat((x0.detach(), d["z"][0, rows], d["v"][0, rows], d["i"][0, rows]), dim=-1).requires_grad_(True)
        gx, gi = RK4().integrate_DAE(x_init=x0, x_func=de_d, i_func=ae_d, t=d["t"][:, rows], x=x0.detach().unsqueeze(0).expand(T, nb, X),
                                     z=d["z"][:, rows], v=d["v"][:, rows], i=d["i"][:, rows], all_initial=a0)
        ((gx * wx[:, rows].to(dev)).sum() + (gi * wi[:, rows].to(dev)).sum()).backward()
        assert _native.last_kernel() == "psn_tc_dae_grad_reduce_kernel"
        return [p.grad.clone().cpu() for p in plist], x0.grad.cpu(), a0.grad.cpu()

    _, full_dx0, full_da0 = gpu_grads(slice(0, B))
    rows = torch.randperm(B, generator=torch.Generator().manual_seed(68))[:24].sort().values
    sub_theta, _, _ = gpu_grads(rows)
    pd = [(Wm.double().requires_grad_(True), bm.double().requires_grad_(True)) for Wm, bm in _params(de_d.x_dot)]
    pa = [(Wm.double().requires_grad_(True), bm.double().requires_grad_(True)) for Wm, bm in _params(ae_d.i_calculator)]
    x064 = host["x0"][rows].double().requires_grad_(True)
    a064 = torch.cat((x064.detach(), host["z"][0, rows].double(), host["v"][0, rows].double(), host["i"][0, rows].double()), dim=-1).requires_grad_(True)
    sx, si = O.integrate_dae("rk4", pd, pa, x064, host["t"][:, rows].double(), x064.detach().unsqueeze(0).expand(T, len(rows), X),
                             host["z"][:, rows].double(), host["v"][:, rows].double(), host["i"][:, rows].double(), a064)
    ((sx * wx[:, rows].double()).sum() + (si * wi[:, rows].double()).sum()).backward()
    ref_theta = [q for pair in pd + pa for q in (pair[0].grad, pair[1].grad)]
    report = [f"theta[{k}] rel {_rel(gv, r):.2e}" for k, (gv, r) in enumerate(zip(sub_theta, ref_theta))]
    report += [f"d_x_init (full-batch rows) rel {_rel(full_dx0[rows], x064.grad):.2e}", f"d_a0 (full-batch rows) rel {_rel(full_da0[rows], a064.grad):.2e}"]
    print("cfg3 full-length gradients vs oracle fp64:\n  " + "\n  ".join(report))
    for gv, r in zip(sub_theta, ref_theta):
        assert _rel(gv, r) <= 3e-5
    assert _rel(full_dx0[rows], x064.grad) <= 3e-5 and _rel(full_da0[rows], a064.grad) <= 3e-5


def test_cfg4_full_size_forward_and_gradients_against_oracle_sample(native_lib, monkeypatch):
    """BASELINE configs[3] per-GPU shard (B = 4096 x 500 RK4 steps, X = Z = H = 128) on the wide tensor-core kernels: forward rows
    against the oracle, latent-input / x0 gradient rows and the sample's parameter gradients against float64 autograd."""
    from oracle import psnode_oracle as O
    from py_psnode_b200 import DE_Func, RK4, _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    dev = "cuda:0"
    torch.manual_seed(71)
    Bw, Nw, Hw = 4096, 500, 128
    Tw = Nw + 1
    de = DE_Func(x_dim=Hw, z_dim=Hw, hidden_dim=Hw, depth=2)
    g = torch.Generator(device=dev).manual_seed(72)
    t = (torch.arange(Tw, dtype=torch.float32, device=dev) * 0.01).view(Tw, 1, 1).repeat(1, Bw, 1)
    z = torch.randn(Tw, Bw, Hw, device=dev, generator=g) * 0.1
    x0 = torch.randn(Bw, Hw, device=dev, generator=g) * 0.1
    w = torch.randn(Tw, Bw, Hw, device=dev, generator=g) * 1e-3
    de_d = de.to(dev)
    plist = list(de_d.parameters())

    def gpu(rows):
        for p in plist:
            p.grad = None
        x0r = x0[rows].clone().requires_grad_(True)
        zr = z[:, rows].clone().requires_grad_(True)
        nb = x0r.shape[0]
        a0 = torch.cat((x0r, zr[0]), dim=-1)
        sol = RK4().integrate_ODE(x_func=de_d, t=t[:, rows], x=x0r.unsqueeze(0).expand(Tw, nb, Hw), z=zr, all_initial=a0)
        assert _native.last_kernel().startswith("psn_wide_fwd_kernel")
        sol.backward(w[:, rows])
        assert _native.last_kernel() == "psn_wide_assemble_kernel"
        return sol.detach().cpu(), [p.grad.clone().cpu() for p in plist], x0r.grad.cpu(), zr.grad.cpu()

    full_sol, _, full_dx0, full_dz = gpu(slice(0, Bw))
    rows = torch.randperm(Bw, generator=torch.Generator().manual_seed(73))[:12].sort().values
    _, sub_theta, _, _ = gpu(rows.to(dev))
    de_c = de_d          # _params() copies the weights to the host; the module itself stays on the GPU
    hx0, hz, ht, hw = x0[rows.to(dev)].cpu(), z[:, rows.to(dev)].cpu(), t[:, rows.to(dev)].cpu(), w[:, rows.to(dev)].cpu()
    want = O.integrate_ode("rk4", _params(de_c.x_dot), ht, hx0.unsqueeze(0).expand(Tw, len(rows), Hw), hz, torch.cat((hx0, hz[0]), dim=-1))
    assert torch.allclose(full_sol[:, rows], want, rtol=RTOL, atol=ATOL), tol_report(full_sol[:, rows], want)
    p64 = [(Wm.double().requires_grad_(True), bm.double().requires_grad_(True)) for Wm, bm in _params(de_c.x_dot)]
    x064, z64 = hx0.double().requires_grad_(True), hz.double().requires_grad_(True)
    sol64 = O.integrate_ode("rk4", p64, ht.double(), x064.unsqueeze(0).expand(Tw, len(rows), Hw), z64, torch.cat((x064, z64[0]), dim=-1))
    (sol64 * hw.double()).sum().backward()
    ref_theta = [q for pair in p64 for q in (pair[0].grad, pair[1].grad)]
    report = [f"theta[{k}] rel {_rel(gv, r):.2e}" for k, (gv, r) in enumerate(zip(sub_theta, ref_theta))]
    report += [f"d_x0 rows rel {_rel(full_dx0[rows], x064.grad):.2e}", f"d_z rows rel {_rel(full_dz[:, rows], z64.grad):.2e}"]
    print("cfg4 full-size gradients vs oracle fp64:\n  " + "\n  ".join(report))
    for gv, r in zip(sub_theta, ref_theta):
        assert _rel(gv, r) <= 2e-5
    assert _rel(full_dx0[rows], x064.grad) <= 2e-5 and _rel(full_dz[:, rows], z64.grad) <= 2e-5


def test_cfg5_shard_gradients_layer_sweep_vs_generic_sweep(native_lib):
    """BASELINE configs[4] per-GPU shard (DAE_02, latent 256, B = 8192 = 64 full trajectory tiles x 2 feature blocks), 80 RK4 steps (two time
    chunks of the hoisted projections, ten turns of the activation ring), one event: every gradient sink of the layer path's tensor-core
    reverse sweep against the CUDA-core generic sweep (itself pinned to float64 autograd at small sizes), and the forward against it too."""
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4, _native
    torch.manual_seed(2025)
    dev = "cuda:0"
    B, N, H = 8192, 80, 256
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(dev)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda: torch.randn(T, B, H, device=dev) * 0.05
    z0, v0 = mk(), mk()
    x_init, i0 = torch.randn(B, H, device=dev) * 0.05, torch.randn(B, H, device=dev) * 0.05
    wx, wi = mk(), mk()
    event_t = t[N // 2].view(B, 1, 1).clone()
    zj0, vj0 = torch.randn(B, 1, H, device=dev) * 0.05, torch.randn(B, 1, H, device=dev) * 0.05
    plist = list(de.parameters()) + list(ae.parameters())
    res = {}
    for impl in ("layer", "generic"):
        for p in plist:
            p.grad = None
        z, v = z0.clone().requires_grad_(True), v0.clone().requires_grad_(True)
        zj, vj = zj0.clone().requires_grad_(True), vj0.clone().requires_grad_(True)
        xi = x_init.clone().requires_grad_(True)
        a0 = torch.cat((x_init, z0[0], v0[0], i0), dim=-1).requires_grad_(True)
        ev = DAE_Event()
        ev.set_event(t=event_t, z=zj, v=vj)
        xs, is_ = RK4(impl=impl).integrate_DAE(x_init=xi, x_func=de, i_func=ae, t=t, x=x_init.unsqueeze(0).expand(T, B, H), z=z, v=v,
                                              i=i0.unsqueeze(0).expand(T, B, H), all_initial=a0, event_fn=ev.event_fn,
                                              jump_change_fn=ev.jump_change_fn)
        ((xs * wx).sum() + (is_ * wi).sum()).backward()
        kern = _native.last_kernel()
        assert kern.startswith("psn_lg_") if impl == "layer" else kern.startswith("psn_g"), kern
        res[impl] = dict(xs=xs.detach(), is_=is_.detach(), theta=[p.grad.clone() for p in plist], z=z.grad, v=v.grad, zj=zj.grad, vj=vj.grad,
                         xi=xi.grad, a0=a0.grad)
    a, b = res["layer"], res["generic"]
    assert torch.allclose(a["xs"], b["xs"], rtol=1e-5, atol=1e-6) and torch.allclose(a["is_"], b["is_"], rtol=1e-5, atol=1e-6)
    pairs = [(f"theta{k}", x, y) for k, (x, y) in enumerate(zip(a["theta"], b["theta"]))]
    pairs += [(k, a[k], b[k]) for k in ("z", "v", "zj", "vj", "xi", "a0")]
    worst = 0.0
    for name, g, r in pairs:
        scale = float(r.abs().max())
        err = float((g - r).abs().max())
        worst = max(worst, err / max(scale, 1e-30))
        assert err <= 3e-5 * scale + 1e-7, f"{name}: err {err:.3e} scale {scale:.3e}"
    print(f"cfg5 shard, layer vs generic sweep: worst relative gradient difference {worst:.2e}")
