#!/usr/bin/env python
"""Golden fixture for the fused encoder / decoder entry (`integrate_DAE_encoded`): the UNMODIFIED reference's
`DAE_Model.forward` (neural_01_DAE_02_direct_encode.py:126-153) at hidden_dim = 128 -- the width the tensor-core layer path
covers -- run on the CPU with Euler and RK4: forward outputs (fp32 and an fp64 restatement through the same code) and, for RK4, the
script's training loss and every parameter gradient of the fp64 restatement (stored as fp32), so that the layer path's reverse sweep is
checked inside the script's own pipeline against the reference's autograd (tests/test_gpu_real_scripts.py).
The ODE counterpart is script_ode02.npz (hidden_dim = 128 already, make_script_golden.py).

    python tests/golden/make_encoded_golden.py        # build container only (needs oracle/_ref or /root/reference)
"""
import copy
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, HERE)


def main():
    import numpy as np
    import torch
    import ref_runner
    import torch.nn.functional as F
    from make_script_golden import make_inputs, script_loss
    torch.set_num_threads(1)
    nd, importlib = ref_runner.load_reference()
    mod = importlib.import_module("neural_01_DAE_02_direct_encode")
    kw = dict(x_dim=4, z_dim=1, v_dim=2, i_dim=2, hidden_dim=128)
    B, T, E = 5, 11, 1
    torch.manual_seed(321)
    base = mod.DAE_Model(**kw)
    d = make_inputs("dae02", kw, B, T, E, torch, seed=19)
    out = {f"in_{k}": v.numpy() for k, v in d.items()}
    for k, v in base.state_dict().items():
        out[f"w_{k}"] = v.detach().numpy().copy()
    out["kw_keys"] = np.array(list(kw.keys()))
    out["kw_vals"] = np.array(list(kw.values()), dtype=np.int64)
    call = lambda m, dd: m.forward(t=dd["t"], x=dd["x"], z=dd["z"], v=dd["v"], i=dd["i"], event_t=dd["event_t"], z_jump=dd["z_jump"],
                                   v_jump=dd["v_jump"])
    for sname, S in (("euler", nd.Euler), ("rk4", nd.RK4)):
        model = copy.deepcopy(base)
        model.solver = S()
        with torch.no_grad():
            preds = call(model, d)
            m64 = copy.deepcopy(base).double()
            m64.solver = S()
            preds64 = call(m64, {k: v.double() for k, v in d.items()})
        for k in range(2):                        # x_pred, i_pred (the reconstruction outputs do not touch the solver)
            out[f"{sname}_pred{k}"] = preds[k].numpy().copy()
            out[f"{sname}_pred64_{k}"] = preds64[k].numpy().copy()
    # training loss + gradients of the fp64 restatement (the arbiter), RK4
    m64 = copy.deepcopy(base).double()
    m64.solver = nd.RK4()
    loss64, _ = script_loss("dae02", m64, {k: v.double() for k, v in d.items()}, F)
    loss64.backward()
    out["rk4_loss64"] = np.array(loss64.item(), dtype=np.float64)
    for k, p in m64.named_parameters():
        if p.grad is not None:
            out[f"rk4_g64_{k}"] = p.grad.detach().numpy().astype(np.float32)
    path = os.path.join(HERE, "script_dae02_h128.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
