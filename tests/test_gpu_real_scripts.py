"""The four reference training scripts' OWN model classes (`ODE_Model` / `DAE_Model` with their encoders, decoders and
`Init_Func`), imported unmodified from oracle/_ref/src, run on the GPU through the repo's `neural_dae` drop-in shim and are
compared with goldens the unmodified reference produced on the CPU (tests/golden/make_script_golden.py): forward outputs at
rtol=1e-5 / atol=1e-6, the script's loss, every parameter gradient (against the reference's fp64 restatement, with the
reference's own fp32-vs-fp64 error as the yardstick) and the weights after one Adam step.

This is the "scripts drop in unchanged" claim executed: the script files are byte-for-byte the reference's
(neural_00_ODE_01_no_encode.py:71-91, neural_00_ODE_02_direct_encode.py:60-89, neural_01_DAE_01_no_encode.py:86-115,
neural_01_DAE_02_direct_encode.py:103-153); only `neural_dae` and `utils` resolve to this repo."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from helpers import ATOL, GOLDEN_DIR, RTOL, tol_report

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "oracle", "_ref", "src")
REF_STUBS = os.path.join(ROOT, "oracle", "_ref", "stubs")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

MODULES = {"ode01": "neural_00_ODE_01_no_encode", "ode02": "neural_00_ODE_02_direct_encode",
           "dae01": "neural_01_DAE_01_no_encode", "dae02": "neural_01_DAE_02_direct_encode"}


@pytest.fixture(scope="module")
def scripts(native_lib):
    if not os.path.isfile(os.path.join(REF_SRC, MODULES["ode01"] + ".py")):
        pytest.skip("oracle/_ref is absent (python oracle/make_ref.py vendors the reference in the build container)")
    import neural_dae                                   # the repo's shim must be the one the scripts bind to
    assert os.path.abspath(neural_dae.__file__).startswith(ROOT) and "_ref" not in neural_dae.__file__
    for p in (REF_SRC, REF_STUBS):                      # appended: repo root (shim `neural_dae`, `utils`) stays in front
        if p not in sys.path:
            sys.path.append(p)
    return {k: importlib.import_module(v) for k, v in MODULES.items()}


@pytest.mark.parametrize("solver", ["euler", "rk4"])
@pytest.mark.parametrize("name", ["ode01", "ode02", "dae01", "dae02"])
def test_real_script_model_on_gpu(scripts, name, solver):
    import neural_dae
    from make_script_golden import LR, script_loss
    from py_psnode_b200 import _native
    mod = scripts[name]
    assert mod.Euler is neural_dae.Euler, "the script must have bound the repo's solver classes"
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"script_{name}.npz"), allow_pickle=False))
    kw = {str(k): int(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
    dev = torch.device("cuda:0")
    Model = mod.ODE_Model if name.startswith("ode") else mod.DAE_Model
    model = Model(**kw)
    model.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")})
    model = model.to(dev)
    model.solver = {"euler": neural_dae.Euler, "rk4": neural_dae.RK4}[solver]()
    d = {k[3:]: torch.from_numpy(v).to(dev) for k, v in g.items() if k.startswith("in_")}
    opt = torch.optim.Adam(model.parameters(), lr=LR)
    model.train()
    loss, preds = script_loss(name, model, d, torch.nn.functional)
    fwd_kernel = _native.last_kernel()
    opt.zero_grad()
    loss.backward()
    # ---- forward outputs and loss ----
    for k, pr in enumerate(preds):
        want = torch.from_numpy(g[f"{solver}_pred{k}"])
        want64 = torch.from_numpy(g[f"{solver}_pred64_{k}"])
        got = pr.detach().cpu()
        assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), f"{name}/{solver} output {k} ({fwd_kernel}): " + tol_report(got, want, want64)
    assert abs(loss.item() - float(g[f"{solver}_loss"])) <= 1e-5 * abs(float(g[f"{solver}_loss"])) + 1e-7
    # ---- gradients: fp64 restatement is the arbiter, the reference's own fp32 error the yardstick ----
    report = []
    for pname, p in model.named_parameters():
        key = f"{solver}_g64_{pname}"
        if key not in g:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, pname
            continue
        g64 = torch.from_numpy(g[key])
        gref = torch.from_numpy(g[f"{solver}_g_{pname}"]).double()
        got = p.grad.detach().cpu().double()
        scale = float(g64.abs().max())
        err, ref_err = float((got - g64).abs().max()), float((gref - g64).abs().max())
        report.append(f"{pname}: err {err:.2e} (reference fp32 {ref_err:.2e}) scale {scale:.2e}")
        assert err <= max(8 * ref_err, 1e-5 * scale) + 1e-9, f"{name}/{solver} grad " + report[-1]
    print(f"{name}/{solver} [{fwd_kernel}]\n  " + "\n  ".join(report))
    # ---- one Adam step (lr 0.005): first-step update = lr * g / (|g| + eps), compared where the gradient is not ~0 ----
    opt.step()
    for pname, p in model.state_dict().items():
        after = torch.from_numpy(g[f"{solver}_after_{pname}"])
        before = torch.from_numpy(g[f"w_{pname}"])
        gkey = f"{solver}_g_{pname}"
        got = p.detach().cpu()
        if gkey not in g:
            assert torch.equal(got, before), pname
            continue
        big = torch.from_numpy(np.abs(g[gkey]) > 1e-6)
        assert torch.allclose(got[big], after[big], rtol=0, atol=2e-3 * LR), f"{name}/{solver} weights after Adam step: {pname}"
        assert float((got - before).abs().max()) <= 1.001 * LR + 1e-9


def test_real_dae02_model_hidden128_trains_on_the_layer_path(scripts):
    """The DAE_02 script's own `DAE_Model` at hidden_dim = 128 (the width of the tensor-core layer path; its `--hidden` default): forward,
    the script's loss and EVERY parameter gradient -- through init_func, the five encoders / decoders, de_func and ae_func, i.e. the
    layer path's reverse sweep incl. its latent-input, jump, x_init and all_initial gradients chained into torch autograd -- against the
    float64 restatement the unmodified reference produced (tests/golden/make_encoded_golden.py)."""
    import neural_dae
    from make_script_golden import script_loss
    from py_psnode_b200 import _native
    mod = scripts["dae02"]
    g = dict(np.load(os.path.join(GOLDEN_DIR, "script_dae02_h128.npz"), allow_pickle=False))
    kw = {str(k): int(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
    dev = torch.device("cuda:0")
    model = mod.DAE_Model(**kw)
    model.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")})
    model = model.to(dev)
    model.solver = neural_dae.RK4()
    d = {k[3:]: torch.from_numpy(v).to(dev) for k, v in g.items() if k.startswith("in_")}
    model.train()
    loss, preds = script_loss("dae02", model, d, torch.nn.functional)
    fwd_kernel = _native.last_kernel()
    assert fwd_kernel.startswith("psn_lg_gemm_kernel"), fwd_kernel
    loss.backward()
    assert _native.last_kernel().startswith("psn_lg_"), _native.last_kernel()
    for k in range(2):
        want, want64 = torch.from_numpy(g[f"rk4_pred{k}"]), torch.from_numpy(g[f"rk4_pred64_{k}"])
        got = preds[k].detach().cpu()
        assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), f"output {k}: " + tol_report(got, want, want64)
    assert abs(loss.item() - float(g["rk4_loss64"])) <= 1e-5 * abs(float(g["rk4_loss64"])) + 1e-7
    report, bad = [], []
    for pname, p in model.named_parameters():
        key = f"rk4_g64_{pname}"
        if key not in g:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, pname
            continue
        g64 = torch.from_numpy(g[key]).double()
        got = p.grad.detach().cpu().double()
        scale, err = float(g64.abs().max()), float((got - g64).abs().max())
        report.append(f"{pname}: err {err:.2e} scale {scale:.2e} rel {err / max(scale, 1e-30):.1e}")
        if not err <= 2e-5 * scale + 1e-8:
            bad.append(report[-1])
    print("dae02 hidden 128 / rk4 [" + fwd_kernel + "]\n  " + "\n  ".join(report))
    assert not bad, bad
