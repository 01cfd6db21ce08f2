"""`Init_Func` + `all_initial` construction in one launch (C ABI `psnode_init_state` / `_backward`,
`solver.init_state`): the first two lines of `DAE_Model.forward` (neural_01_DAE_01_no_encode.py:50-58, :98-99) against torch -- values,
and every gradient against float64 autograd -- and inside the DAE_01 script's own pipeline against the reference golden."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import ATOL, GOLDEN_DIR, RTOL, tol_report

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "oracle", "_ref", "src")
REF_STUBS = os.path.join(ROOT, "oracle", "_ref", "stubs")
DEV = "cuda:0"


class InitFunc(nn.Module):
    """Shape of the scripts' Init_Func (neural_01_DAE_01_no_encode.py:50-58)."""

    def __init__(self, x_dim, z_dim, v_dim, i_dim, hidden_dim):
        super().__init__()
        self.init_fun = nn.Sequential(nn.Linear(z_dim + v_dim + i_dim, hidden_dim), nn.ELU(), nn.Linear(hidden_dim, hidden_dim), nn.ELU(),
                                      nn.Linear(hidden_dim, x_dim))

    def forward(self, z0, v0, i0):
        return self.init_fun(torch.cat([z0, v0, i0], dim=-1))


@pytest.mark.parametrize("B,X,Z,V,I,H", [(37, 16, 1, 2, 4, 64), (4096, 16, 1, 2, 4, 64), (50, 32, 1, 2, 2, 128), (9, 5, 0, 3, 2, 32)])
def test_init_state_values_and_gradients(native_lib, B, X, Z, V, I, H):
    from py_psnode_b200 import RK4, _native
    torch.manual_seed(7 + B)
    T = 5
    f = InitFunc(X, Z, V, I, H)
    f64 = InitFunc(X, Z, V, I, H).double()
    f64.load_state_dict({k: v.double() for k, v in f.state_dict().items()})
    series = {k: torch.randn(B, T, w) for k, w in (("z", Z), ("v", V), ("i", I))}        # batch-major storage, as the scripts hold it
    w1, w2 = torch.randn(B, X), torch.randn(B, X + Z + V + I)
    # float64 autograd
    r64 = {k: s.double().permute(1, 0, 2)[0].clone().requires_grad_(True) for k, s in series.items()}
    x64 = f64(r64["z"], r64["v"], r64["i"])
    a64 = torch.cat((x64, r64["z"], r64["v"], r64["i"]), dim=-1)
    ((x64 * w1.double()).sum() + (a64 * w2.double()).sum()).backward()
    # one launch
    fd = f.to(DEV)
    full = {k: s.to(DEV).requires_grad_(True) for k, s in series.items()}
    rows = {k: s.permute(1, 0, 2)[0] for k, s in full.items()}                          # strided views of the first grid row
    x0, a0 = RK4.init_state(fd, rows["z"] if Z else None, rows["v"], rows["i"])
    assert _native.last_kernel() == "psn_init_state_kernel"
    assert torch.allclose(x0.cpu().double(), x64.detach(), rtol=2e-6, atol=2e-6), tol_report(x0.cpu(), x64.detach())
    assert torch.allclose(a0.cpu().double(), a64.detach(), rtol=2e-6, atol=2e-6)
    ((x0 * w1.to(DEV)).sum() + (a0 * w2.to(DEV)).sum()).backward()
    assert _native.last_kernel() == "psn_init_reduce_kernel"
    pairs = [(n, p.grad, dict(f64.named_parameters())[n].grad) for n, p in fd.named_parameters()]
    pairs += [(k, full[k].grad[:, 0], r64[k].grad) for k in ("z", "v", "i") if series[k].shape[-1]]
    for name, g, g64 in pairs:
        scale = float(g64.abs().max())
        err = float((g.cpu().double() - g64).abs().max())
        assert err <= 5e-6 * scale + 1e-7, f"{name}: err {err:.3e} scale {scale:.3e}"
    for k in ("z", "v", "i"):
        if series[k].shape[-1]:
            assert float(full[k].grad[:, 1:].abs().max()) == 0.0          # only the first grid row is an input


@pytest.mark.parametrize("solver", ["euler", "rk4"])
def test_init_state_in_the_dae01_script_pipeline(native_lib, solver):
    """DAE_Model.forward of the DAE_01 script with its first two lines replaced by solver.init_state: outputs equal to the golden
    the unmodified reference produced (same bar as tests/test_gpu_real_scripts.py)."""
    if not os.path.isfile(os.path.join(REF_SRC, "neural_01_DAE_01_no_encode.py")):
        pytest.skip("oracle/_ref is absent (python oracle/make_ref.py vendors the reference in the build container)")
    import neural_dae
    for p in (REF_SRC, REF_STUBS):
        if p not in sys.path:
            sys.path.append(p)
    mod = importlib.import_module("neural_01_DAE_01_no_encode")
    g = dict(np.load(os.path.join(GOLDEN_DIR, "script_dae01.npz"), allow_pickle=False))
    kw = {str(k): int(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
    model = mod.DAE_Model(**kw)
    model.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")})
    model = model.to(DEV)
    d = {k[3:]: torch.from_numpy(v).to(DEV) for k, v in g.items() if k.startswith("in_")}
    S = {"euler": neural_dae.Euler, "rk4": neural_dae.RK4}[solver]()
    tm = lambda q: q.permute(1, 0, 2)
    with torch.no_grad():
        model.event.set_event(t=d["event_t"], z=d["z_jump"], v=d["v_jump"])
        x0, a0 = S.init_state(model.init_func, tm(d["z"])[0], tm(d["v"])[0], tm(d["i"])[0])
        xs, is_ = S.integrate_DAE(x_init=x0, x_func=model.de_func, i_func=model.ae_func, t=tm(d["t"]), x=tm(d["x"]), z=tm(d["z"]), v=tm(d["v"]),
                                  i=tm(d["i"]), all_initial=a0, event_fn=model.event.event_fn, jump_change_fn=model.event.jump_change_fn)
    for k, got in enumerate((xs, is_)):
        got = got.permute(1, 0, 2).cpu()
        want, want64 = torch.from_numpy(g[f"{solver}_pred{k}"]), torch.from_numpy(g[f"{solver}_pred64_{k}"])
        assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), f"output {k}: " + tol_report(got, want, want64)
