"""GPU parity of the reverse sweep (discrete adjoint, C ABI `psnode_backward`) against the reference's own autograd
gradients stored in tests/golden/*.npz (fp32 `g_*` and fp64 `g64_*`, produced by the unmodified reference).

Tolerance: gradients are sums over B x T x stages terms, so two correct fp32 implementations differ by summation order.
Each tensor must match the fp64 reference to rtol=1e-5 * max|g64| + atol 1e-7 OR within 8x of the error the reference's
own fp32 autograd has against fp64 -- whichever is looser -- and elementwise allclose(rtol=1e-3, atol=1e-5*max|g64|)."""
import numpy as np
import pytest
import torch

from helpers import AE, DE, golden_names, load_golden, params_of, tm

pytestmark = pytest.mark.gpu


def _solver(name, impl):
    from py_psnode_b200 import Euler, Midpoint, RK4
    return {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[name](impl=impl)


def _leaf(a, dev):
    return torch.from_numpy(a).to(dev).requires_grad_(True)


def run_with_grads(d, name, impl="auto", dev="cuda:0"):
    """Mirror of make_golden.py's `forward(..., want_grads=True)` on the CUDA path.  Returns {key: grad (cpu numpy)}."""
    from py_psnode_b200 import DAE_Event, ODE_Event
    dae = str(d["kind"]) == "dae"
    input_grads = ("g_z" in d) or ("g_x" in d)
    leaves = {}
    de = DE(params_of(d, "de")).to(dev)
    ae = AE(params_of(d, "ae")).to(dev) if dae else None
    t_bt = torch.from_numpy(d["t"]).to(dev)
    series = {}
    for key in ("x", "z", "v", "i"):
        if key in d:
            ten = torch.from_numpy(d[key]).to(dev)
            if input_grads:
                ten.requires_grad_(True)
                leaves[key] = ten
            series[key] = ten
    has_event = "noevent" not in name
    zj = torch.from_numpy(d["z_jump"]).to(dev)
    vj = torch.from_numpy(d["v_jump"]).to(dev) if dae else None
    if input_grads and has_event:
        zj.requires_grad_(True); leaves["z_jump"] = zj
        if dae:
            vj.requires_grad_(True); leaves["v_jump"] = vj
    event_fn = jump_fn = None
    if has_event:
        ev = DAE_Event() if dae else ODE_Event()
        if dae:
            ev.set_event(t=torch.from_numpy(d["event_t"]).to(dev), z=zj, v=vj)
        else:
            ev.set_event(t=torch.from_numpy(d["event_t"]).to(dev), z=zj)
        event_fn, jump_fn = ev.event_fn, ev.jump_change_fn
    P = lambda q: q.permute(1, 0, 2)
    solver = _solver(str(d["solver"]), impl)
    if dae:
        xi = _leaf(d["x_init"], dev)
        leaves["x_init"] = xi
        a0 = torch.cat((xi, P(series["z"])[0], P(series["v"])[0], P(series["i"])[0]), dim=-1)
        if not input_grads:
            a0 = a0.detach().clone().requires_grad_(True)
            leaves["all_initial"] = a0
        xs, is_ = solver.integrate_DAE(x_init=xi, x_func=de, i_func=ae, t=P(t_bt), x=P(series["x"]), z=P(series["z"]),
                                       v=P(series["v"]), i=P(series["i"]), all_initial=a0, event_fn=event_fn,
                                       jump_change_fn=jump_fn, input_true_x=bool(d["teacher_x"]), input_true_i=bool(d["teacher_i"]))
        loss = (xs * torch.from_numpy(d["gx"]).to(dev)).sum() + (is_ * torch.from_numpy(d["gi"]).to(dev)).sum()
    else:
        a0 = torch.cat((P(series["x"])[0], P(series["z"])[0]), dim=-1)
        if not input_grads:
            a0 = a0.detach().clone().requires_grad_(True)
            leaves["all_initial"] = a0
        xs = solver.integrate_ODE(x_func=de, t=P(t_bt), x=P(series["x"]), z=P(series["z"]), all_initial=a0, event_fn=event_fn,
                                  jump_change_fn=jump_fn, input_true_x=bool(d["teacher_x"]))
        loss = (xs * torch.from_numpy(d["gx"]).to(dev)).sum()
    loss.backward()
    out = {}
    for net, mod in (("de", de.x_dot), ("ae", ae.i_calculator if dae else [])):
        k = 0
        for m in mod:
            if isinstance(m, torch.nn.Linear):
                out[f"{net}_W{k}"] = m.weight.grad.cpu().numpy()
                out[f"{net}_b{k}"] = m.bias.grad.cpu().numpy()
                k += 1
    for key, ten in leaves.items():
        out[key] = (ten.grad if ten.grad is not None else torch.zeros_like(ten)).cpu().numpy()
    return out


def grad_errors(d, got):
    rows = []
    for key in sorted(k[4:] for k in d if k.startswith("g64_")):
        g64, g32 = d["g64_" + key], d["g_" + key]
        if key not in got or g64.size == 0:
            continue
        mine = got[key].reshape(g64.shape).astype(np.float64)
        scale = float(np.abs(g64).max())
        rows.append((key, scale, float(np.abs(mine - g64).max()), float(np.abs(g32 - g64).max()), mine, g64))
    return rows


CASES = [n for n in golden_names() if "model" not in n and "long" not in n]


def _tape_sweep_expected(d):
    """Name of the reduce kernel of the tensor-core reverse sweep (activation tape) when it must be the one that runs --
    H = 64 4-layer nets, X = 16, no teacher forcing, no input-series gradients -- else None."""
    if ("g_z" in d) or ("g_x" in d) or bool(d["teacher_x"]) or ("teacher_i" in d and bool(d["teacher_i"])):
        return None
    de = [d[f"de_W{k}"].shape for k in range(8) if f"de_W{k}" in d]
    X = 16
    if str(d["kind"]) == "ode":
        S = d["x"].shape[-1] + d["z"].shape[-1]
        ok = d["x"].shape[-1] == X and d["z"].shape[-1] <= 8 and de == [(64, 3 * S), (64, 64), (64, 64), (16, 64)]
        return "psn_tc_grad_reduce_kernel" if ok else None
    Z, V, I = d["z"].shape[-1], d["v"].shape[-1], d["i"].shape[-1]
    S = X + Z + V + I
    ae = [d[f"ae_W{k}"].shape for k in range(8) if f"ae_W{k}" in d]
    ok = (d["x_init"].shape[-1] == X and Z + V + I <= 8 and de == [(64, 3 * S), (64, 64), (64, 64), (16, 64)]
          and ae == [(64, S + X + Z + V), (64, 64), (64, 64), (I, 64)])
    return "psn_tc_dae_grad_reduce_kernel" if ok else None


@pytest.mark.parametrize("sweep", ["tape", "recompute"])
@pytest.mark.parametrize("name", CASES)
def test_backward_matches_reference(native_lib, name, sweep, monkeypatch):
    """sweep = tape: the forward records the activation tape and the tensor-core reverse sweep consumes it (where the
    problem supports it); sweep = recompute: PSNODE_TAPE_MAX_GB=0 forces the generic recomputing sweep everywhere."""
    from py_psnode_b200 import _native
    d = load_golden(name)
    if "gx" not in d:
        pytest.skip("fixture has no gradients")
    if sweep == "recompute":
        monkeypatch.setenv("PSNODE_TAPE_MAX_GB", "0")
    else:
        monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    got = run_with_grads(d, name)
    if sweep == "tape" and _tape_sweep_expected(d):
        assert _native.last_kernel() == _tape_sweep_expected(d), _native.last_kernel()
    else:
        assert _native.last_kernel() == "psn_grad_reduce_kernel", _native.last_kernel()
    rows = grad_errors(d, got)
    assert rows, "no gradient tensors compared"
    want_keys = {k[4:] for k in d if k.startswith("g64_") and d[k].size}
    assert want_keys <= set(got), f"missing gradients: {sorted(want_keys - set(got))}"
    bad = []
    for key, scale, err, ref_err, mine, g64 in rows:
        bound = max(8.0 * ref_err, 1e-5 * scale + 1e-7)
        if err > bound or not np.allclose(mine, g64, rtol=1e-3, atol=1e-5 * scale + 1e-7):
            bad.append(f"{key}: max|got-g64|={err:.3e} (ref fp32 err {ref_err:.3e}, scale {scale:.3e}, bound {bound:.3e})")
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("impl", ["tc", "tc8"])
@pytest.mark.parametrize("name", ["ode01_rk4_small", "ode01_midpoint_small", "ode01_rk4_noevent"])
def test_backward_tape_written_by_every_forward_kernel(native_lib, name, impl, monkeypatch):
    """Both tensor-core forward kernels record the same tape layout: tensor-core reverse sweep on a tape written by each."""
    from py_psnode_b200 import _native
    monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
    d = load_golden(name)
    got = run_with_grads(d, name, impl=impl)
    assert _native.last_kernel() == "psn_tc_grad_reduce_kernel", _native.last_kernel()
    for key, scale, err, ref_err, mine, g64 in grad_errors(d, got):
        assert err <= max(8.0 * ref_err, 1e-5 * scale + 1e-7), f"{key}: {err:.3e} vs ref {ref_err:.3e} scale {scale:.3e}"


def test_model_loss_gradients(native_lib):
    """Reference ODE_Model.forward (permuted views) + masked-MSE loss: parameter gradients of a real training step."""
    from py_psnode_b200 import ODE_Event, RK4
    d = load_golden("ode01_model_rk4")
    dev = "cuda:0"
    de = DE(params_of(d, "de")).to(dev)
    t, x, z = (torch.from_numpy(d[k]).to(dev) for k in ("t", "x", "z"))
    ev = ODE_Event()
    ev.set_event(t=torch.from_numpy(d["event_t"]).to(dev), z=torch.from_numpy(d["z_jump"]).to(dev))
    a0 = torch.cat((x[:, 0], z[:, 0]), dim=-1)
    sol = RK4().integrate_ODE(x_func=de, t=t.permute(1, 0, 2), x=x.permute(1, 0, 2), z=z.permute(1, 0, 2), all_initial=a0,
                              event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
    pred = sol.permute(1, 0, 2)
    mask = torch.from_numpy(d["mask"]).to(dev)
    loss = torch.sum(torch.nn.functional.mse_loss(pred, x, reduction="none") * mask) / torch.sum(mask)
    np.testing.assert_allclose(loss.item(), float(d["loss"]), rtol=1e-5)
    loss.backward()
    k = 0
    for m in de.x_dot:
        if isinstance(m, torch.nn.Linear):
            np.testing.assert_allclose(m.weight.grad.cpu().numpy(), d[f"g_de_W{k}"], rtol=2e-4, atol=1e-7)
            np.testing.assert_allclose(m.bias.grad.cpu().numpy(), d[f"g_de_b{k}"], rtol=2e-4, atol=1e-7)
            k += 1


if __name__ == "__main__":      # error table for tolerance tuning: python tests/test_gpu_backward.py
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for name in CASES:
        d = load_golden(name)
        if "gx" not in d:
            continue
        try:
            got = run_with_grads(d, name)
        except Exception as exc:       # noqa: BLE001
            print(f"{name}: EXC {type(exc).__name__}: {exc}")
            continue
        for key, scale, err, ref_err, _, _ in grad_errors(d, got):
            flag = "" if err <= max(8 * ref_err, 1e-5 * scale + 1e-7) else "  <<<<"
            print(f"{name:28s} {key:12s} scale {scale:9.3e} err {err:9.3e} ref32 {ref_err:9.3e}{flag}")


def test_tape_sweep_in_batch_chunks_when_the_tape_does_not_fit(native_lib, monkeypatch):
    """B = 200 trajectories with room for a 32-trajectory tape only: the reverse sweep re-integrates the batch in 7 chunks on
    the tensor-core kernels (not the generic sweep) and must reproduce the gradients of the single-tape run."""
    from py_psnode_b200 import DE_Func, ODE_Event, RK4, _native
    torch.manual_seed(61)
    dev = "cuda:0"
    B, N, X, Z, H = 200, 50, 16, 2, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    x = torch.randn(T, B, X, device=dev) * 0.1
    z = torch.randn(T, B, Z, device=dev) * 0.1
    w = torch.randn(T, B, X, device=dev) * 0.1
    ev = ODE_Event()
    ev.set_event(t=t[N // 2].view(B, 1, 1).clone(), z=torch.randn(B, 1, Z, device=dev) * 0.1)
    per_group_bytes = N * 4 * 3328 * 4                      # (T-1) steps x 4 stages x 3328 floats
    grads = {}
    for mode in ("single", "chunked"):
        if mode == "chunked":
            monkeypatch.setenv("PSNODE_TAPE_MAX_GB", str(2.5 * per_group_bytes / 2 ** 30))      # room for two 16-trajectory groups
        else:
            monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
        for p in de.parameters():
            p.grad = None
        xd = x.clone().requires_grad_(True)
        a0 = torch.cat((x[0], z[0]), dim=-1).requires_grad_(True)
        n0 = _native.launch_count()
        sol = RK4().integrate_ODE(x_func=de, t=t, x=xd, z=z, all_initial=a0, event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
        (sol * w).sum().backward()
        assert _native.last_kernel() == "psn_tc_grad_reduce_kernel", _native.last_kernel()
        launches = _native.launch_count() - n0
        assert launches > 20 if mode == "chunked" else launches < 10, launches
        grads[mode] = [p.grad.clone() for p in de.parameters()] + [xd.grad.clone(), a0.grad.clone()]
    for k, (a, b) in enumerate(zip(grads["chunked"], grads["single"])):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 2e-6 * scale + 1e-9, f"tensor {k}"


def test_dae_tape_sweep_in_batch_chunks(native_lib, monkeypatch):
    """Same as above for the DAE sweep (B = 72 trajectories, tape room for 32): chunked == single-tape gradients."""
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, RK4, _native
    torch.manual_seed(62)
    dev = "cuda:0"
    B, N, X, Z, V, I, H = 72, 20, 16, 1, 2, 4, 64
    T = N + 1
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I).to(dev)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z).to(dev)
    t = (torch.arange(T, dtype=torch.float32, device=dev) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    mk = lambda wd: torch.randn(T, B, wd, device=dev) * 0.1
    x, z, v, i = mk(X), mk(Z), mk(V), mk(I)
    x_init = torch.randn(B, X, device=dev) * 0.1
    wx, wi = mk(X), mk(I)
    ev = DAE_Event()
    ev.set_event(t=t[N // 2].view(B, 1, 1).clone(), z=torch.randn(B, 1, Z, device=dev) * 0.1, v=torch.randn(B, 1, V, device=dev) * 0.1)
    per_group_bytes = (N * 5 + 2) * 3328 * 4
    plist = list(de.parameters()) + list(ae.parameters())
    grads = {}
    for mode in ("single", "chunked"):
        if mode == "chunked":
            monkeypatch.setenv("PSNODE_TAPE_MAX_GB", str(2.5 * per_group_bytes / 2 ** 30))
        else:
            monkeypatch.delenv("PSNODE_TAPE_MAX_GB", raising=False)
        for p in plist:
            p.grad = None
        xi = x_init.clone().requires_grad_(True)
        a0 = torch.cat((x_init, z[0], v[0], i[0]), dim=-1).requires_grad_(True)
        n0 = _native.launch_count()
        gx, gi = RK4().integrate_DAE(x_init=xi, x_func=de, i_func=ae, t=t, x=x, z=z, v=v, i=i, all_initial=a0, event_fn=ev.event_fn,
                                     jump_change_fn=ev.jump_change_fn)
        ((gx * wx).sum() + (gi * wi).sum()).backward()
        assert _native.last_kernel() == "psn_tc_dae_grad_reduce_kernel", _native.last_kernel()
        launches = _native.launch_count() - n0
        assert launches > 10 if mode == "chunked" else launches < 10, launches
        grads[mode] = [p.grad.clone() for p in plist] + [xi.grad.clone(), a0.grad.clone()]
    for k, (a, b) in enumerate(zip(grads["chunked"], grads["single"])):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 2e-6 * scale + 1e-9, f"tensor {k}"
