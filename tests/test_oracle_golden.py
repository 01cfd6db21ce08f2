"""Pin the oracle: oracle/psnode_oracle.py must reproduce every output of the unmodified reference stored in
tests/golden/ (forward fp32, forward fp64 and every autograd gradient).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import golden_names, load_golden, params_of, tm
from oracle import psnode_oracle as O

# The fp32 oracle runs the same ATen kernels in the same order as the reference, so on the host that generated the
# fixtures it is bit-exact; the tolerance only absorbs a different CPU picking a different GEMM micro-kernel.
F32 = dict(rtol=2e-6, atol=2e-7)
F64 = dict(rtol=1e-12, atol=1e-13)


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _events(d, name, dtype, dae):
    if "noevent" in name:
        return (None, None, None) if dae else (None, None)
    ev = torch.from_numpy(d["event_t"]).to(dtype)
    if dae:
        return ev, torch.from_numpy(d["z_jump"]).to(dtype), torch.from_numpy(d["v_jump"]).to(dtype)
    return ev, torch.from_numpy(d["z_jump"]).to(dtype)


def _ode(d, name, dtype, leaves=None):
    de = params_of(d, "de", dtype)
    t, x, z = tm(d["t"], dtype=dtype), tm(d["x"], dtype=dtype), tm(d["z"], dtype=dtype)
    ev, zj = _events(d, name, dtype, False)
    if leaves is not None:
        for W, b in de:
            W.requires_grad_(True); b.requires_grad_(True)
        leaves["de"] = de
        if "g_x" in d:
            x = x.detach().clone().requires_grad_(True); z = z.detach().clone().requires_grad_(True)
            zj = zj.detach().clone().requires_grad_(True)
            leaves.update(x=x, z=z, z_jump=zj)
    a0 = torch.cat((x[0], z[0]), dim=-1)
    if leaves is not None and "g_all_initial" in d:
        a0 = a0.detach().clone().requires_grad_(True)
        leaves["all_initial"] = a0
    return O.integrate_ode(str(d["solver"]), de, t, x, z, a0, ev, zj, teacher_x=bool(d["teacher_x"]))


def _dae(d, name, dtype, leaves=None):
    de, ae = params_of(d, "de", dtype), params_of(d, "ae", dtype)
    t, x, z, v, i = (tm(d[k], dtype=dtype) for k in ("t", "x", "z", "v", "i"))
    ev, zj, vj = _events(d, name, dtype, True)
    xi = torch.from_numpy(d["x_init"]).to(dtype)
    if leaves is not None:
        for W, b in de + ae:
            W.requires_grad_(True); b.requires_grad_(True)
        leaves["de"], leaves["ae"] = de, ae
        xi = xi.requires_grad_(True); leaves["x_init"] = xi
        if "g_z" in d:
            z, v, x, i = (q.detach().clone().requires_grad_(True) for q in (z, v, x, i))
            zj = zj.detach().clone().requires_grad_(True); vj = vj.detach().clone().requires_grad_(True)
            leaves.update(z=z, v=v, x=x, i=i, z_jump=zj, v_jump=vj)
    a0 = torch.cat((xi, z[0], v[0], i[0]), dim=-1)
    if leaves is not None and "g_all_initial" in d:
        a0 = a0.detach().clone().requires_grad_(True)
        leaves["all_initial"] = a0
    return O.integrate_dae(str(d["solver"]), de, ae, xi, t, x, z, v, i, a0, ev, zj, vj,
                           teacher_x=bool(d["teacher_x"]), teacher_i=bool(d["teacher_i"]))


@pytest.mark.parametrize("name", [n for n in golden_names("ode0") if "model" not in n])
def test_oracle_ode_forward(name):
    d = load_golden(name)
    with torch.no_grad():
        got = _ode(d, name, torch.float32)
        got64 = _ode(d, name, torch.float64)
    np.testing.assert_allclose(got.numpy(), d["x_sol"], **F32)
    np.testing.assert_allclose(got64.numpy(), d["x_sol64"], **F64)


@pytest.mark.parametrize("name", golden_names("dae0"))
def test_oracle_dae_forward(name):
    d = load_golden(name)
    with torch.no_grad():
        gx, gi = _dae(d, name, torch.float32)
        gx64, gi64 = _dae(d, name, torch.float64)
    np.testing.assert_allclose(gx.numpy(), d["x_sol"], **F32)
    np.testing.assert_allclose(gi.numpy(), d["i_sol"], **F32)
    np.testing.assert_allclose(gx64.numpy(), d["x_sol64"], **F64)
    np.testing.assert_allclose(gi64.numpy(), d["i_sol64"], **F64)


def _grad_names(d, tag):
    pre = f"g{tag}_"
    return [k[len(pre):] for k in d if k.startswith(pre)]


def _leaf_grad(leaves, key):
    if key.startswith("de_") or key.startswith("ae_"):
        net, wb = key[:2], key[3:]
        W, b = leaves[net][int(wb[1:])]
        ten = W if wb[0] == "W" else b
    else:
        ten = leaves[key]
    return torch.zeros_like(ten) if ten.grad is None else ten.grad


@pytest.mark.parametrize("name", [n for n in golden_names() if "model" not in n and "long" not in n])
def test_oracle_gradients(name):
    """Autograd through the oracle (fp64) equals autograd through the reference (fp64): pins the backward oracle."""
    d = load_golden(name)
    if "gx" not in d:
        pytest.skip("fixture has no gradients")
    leaves = {}
    dt = torch.float64
    if str(d["kind"]) == "ode":
        sol = _ode(d, name, dt, leaves)
        loss = (sol * torch.from_numpy(d["gx"]).to(dt)).sum()
    else:
        xs, is_ = _dae(d, name, dt, leaves)
        loss = (xs * torch.from_numpy(d["gx"]).to(dt)).sum() + (is_ * torch.from_numpy(d["gi"]).to(dt)).sum()
    loss.backward()
    for key in _grad_names(d, "64"):
        want = d["g64_" + key]
        got = _leaf_grad(leaves, key)
        if key in ("x", "z", "v", "i"):           # stored batch-major; ours are permuted views of batch-major leaves
            got = got.permute(1, 0, 2) if got.shape != want.shape else got
        np.testing.assert_allclose(got.numpy().reshape(want.shape), want, rtol=1e-9, atol=1e-12, err_msg=key)


def test_oracle_model_forward_and_loss():
    """Whole reference ODE_Model.forward + masked-MSE loss + parameter gradients (fixture ode01_model_rk4)."""
    d = load_golden("ode01_model_rk4")
    de = params_of(d, "de")
    for W, b in de:
        W.requires_grad_(True); b.requires_grad_(True)
    t, x, z = tm(d["t"]), tm(d["x"]), tm(d["z"])
    a0 = torch.cat((x[0], z[0]), dim=-1)
    sol = O.integrate_ode("rk4", de, t, x, z, a0, torch.from_numpy(d["event_t"]), torch.from_numpy(d["z_jump"]))
    pred = sol.permute(1, 0, 2)
    np.testing.assert_allclose(pred.detach().numpy(), d["x_pred"], **F32)
    mask = torch.from_numpy(d["mask"])
    xb = torch.from_numpy(d["x"])
    loss = torch.sum(torch.nn.functional.mse_loss(pred, xb, reduction="none") * mask) / torch.sum(mask)
    np.testing.assert_allclose(loss.item(), float(d["loss"]), rtol=1e-5)
    loss.backward()
    for k, (W, b) in enumerate(de):
        np.testing.assert_allclose(W.grad.numpy(), d[f"g_de_W{k}"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(b.grad.numpy(), d[f"g_de_b{k}"], rtol=2e-4, atol=1e-7)


def test_multiple_event_matches_raise():
    t = torch.zeros(3, 2, 1)
    with pytest.raises(ValueError):
        O.event_index(t[0], torch.zeros(2, 2, 1))
