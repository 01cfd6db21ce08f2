#!/usr/bin/env python
"""Golden fixtures of the four reference TRAINING SCRIPTS' own models (`ODE_Model` / `DAE_Model`, incl. encoders, decoders,
`Init_Func`), produced by the UNMODIFIED reference on the CPU:  forward outputs, the script's loss, every parameter gradient
(fp32 and an fp64 restatement through the same code) and the weights after ONE `torch.optim.Adam(lr=0.005)` step
(neural_00_ODE_01_no_encode.py:294,350-360; neural_00_ODE_02_direct_encode.py:208,264-275;
 neural_01_DAE_01_no_encode.py:350,409-424; neural_01_DAE_02_direct_encode.py:296,355-370).

    python tests/golden/make_script_golden.py        # build container only (needs oracle/_ref or /root/reference)

tests/test_gpu_real_scripts.py imports the SAME script files (from oracle/_ref/src) with the repo's `neural_dae` shim first
on sys.path, loads these weights, runs on the GPU and compares.  The script-default solver is Euler (hard-coded in every
model); each case is also stored with `model.solver = RK4()`.
"""
import copy
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle"))

CASES = {
    # name: (module, model kwargs, B, T, events)
    "ode01": ("neural_00_ODE_01_no_encode", dict(x_dim=5, z_dim=2, hidden_dim=64), 6, 14, 1),
    "ode02": ("neural_00_ODE_02_direct_encode", dict(x_dim=3, z_dim=2, hidden_dim=128), 5, 10, 1),
    "dae01": ("neural_01_DAE_01_no_encode", dict(x_dim=6, z_dim=1, v_dim=2, i_dim=2, hidden_dim=64), 6, 12, 1),
    "dae02": ("neural_01_DAE_02_direct_encode", dict(x_dim=4, z_dim=1, v_dim=2, i_dim=2, hidden_dim=32), 4, 8, 1),
}
LR = 0.005


def script_loss(name, model, d, F):
    """The loss of the script's training loop, verbatim formulas."""
    import torch
    mse = F.mse_loss
    if name == "ode01":
        x_pred = model.forward(t=d["t"], x=d["x"], z=d["z"], event_t=d["event_t"], z_jump=d["z_jump"])
        x0_loss = mse(d["x"][:, 0, :], x_pred[:, 0, :]).view(1)
        x_loss = torch.sum(torch.sum(mse(x_pred, d["x"], reduction="none") * d["mask"], dim=1), dim=0) / torch.sum(d["mask"])
        return torch.sum(x0_loss) + torch.sum(x_loss), (x_pred,)
    if name == "ode02":
        x_pred, x_re = model.forward(t=d["t"], x=d["x"], z=d["z"], event_t=d["event_t"], z_jump=d["z_jump"])
        x0_loss = mse(d["x"][:, 0, :], x_pred[:, 0, :]).view(1)
        x_loss = torch.sum(torch.sum(mse(x_pred, d["x"], reduction="none") * d["mask"], dim=1), dim=0) / torch.sum(d["mask"])
        x_recon = mse(x_re, d["x"]).view(1)
        return torch.sum(x0_loss) + torch.sum(x_loss) + torch.sum(x_recon), (x_pred, x_re)
    if name == "dae01":
        x_pred, i_pred = model.forward(t=d["t"], x=d["x"], z=d["z"], v=d["v"], i=d["i"], event_t=d["event_t"], z_jump=d["z_jump"],
                                       v_jump=d["v_jump"])
        x, i, mask = d["x"], d["i"], d["mask"]
        x_loss = (torch.sum(mse(x_pred, x, reduction="none") * mask)
                  + torch.sum(mse(x_pred[:, :, 1:2], x[:, :, 1:2], reduction="none") * mask) * 9) / torch.sum(mask)
        i_loss = torch.sum(mse(i_pred, i, reduction="none") * mask) / torch.sum(mask)
        return x_loss + i_loss + mse(x[:, 0, :], x_pred[:, 0, :]) + mse(i[:, 0, :], i_pred[:, 0, :]), (x_pred, i_pred)
    x_pred, i_pred, x_re, i_re = model.forward(t=d["t"], x=d["x"], z=d["z"], v=d["v"], i=d["i"], event_t=d["event_t"],
                                               z_jump=d["z_jump"], v_jump=d["v_jump"])
    x, i, mask = d["x"], d["i"], d["mask"]
    x_loss = torch.sum(mse(x_pred, x, reduction="none") * mask) / torch.sum(mask)
    i_loss = torch.sum(mse(i_pred, i, reduction="none") * mask) / torch.sum(mask)
    recon = mse(x_re, x) + mse(i_re, i)
    return x_loss + i_loss + mse(x[:, 0, :], x_pred[:, 0, :]) + mse(i[:, 0, :], i_pred[:, 0, :]) + recon, (x_pred, i_pred, x_re, i_re)


def make_inputs(name, kw, B, T, E, torch, seed):
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g) * 0.1
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(1, T, 1).repeat(B, 1, 1).contiguous()
    d = dict(t=t, x=rnd(B, T, kw["x_dim"]), z=rnd(B, T, kw["z_dim"]))
    mask = torch.ones(B, T, 1)
    mask[:, T - 2:, :] = 0.0                                      # the scripts mask the padded tail
    d["mask"] = mask
    steps = [T // 3][:E]
    d["event_t"] = torch.stack([t[:, s, 0] for s in steps], dim=1).view(B, E, 1).clone()
    d["z_jump"] = rnd(B, E, kw["z_dim"])
    if name.startswith("dae"):
        d["v"], d["i"] = rnd(B, T, kw["v_dim"]), rnd(B, T, kw["i_dim"])
        d["v_jump"] = rnd(B, E, kw["v_dim"])
    return d


def main():
    import numpy as np
    import torch
    import torch.nn.functional as F
    import ref_runner
    torch.set_num_threads(1)
    nd, importlib = ref_runner.load_reference()
    for name, (modname, kw, B, T, E) in CASES.items():
        mod = importlib.import_module(modname)
        Model = mod.ODE_Model if name.startswith("ode") else mod.DAE_Model
        torch.manual_seed(100 + len(name) + B)
        base = Model(**kw)
        d = make_inputs(name, kw, B, T, E, torch, seed=7 + B)
        out = {f"in_{k}": v.numpy() for k, v in d.items()}
        for k, v in base.state_dict().items():
            out[f"w_{k}"] = v.detach().numpy().copy()
        out["kw_keys"] = np.array(list(kw.keys()))
        out["kw_vals"] = np.array(list(kw.values()), dtype=np.int64)
        for sname, S in (("euler", nd.Euler), ("rk4", nd.RK4)):
            model = copy.deepcopy(base)
            model.solver = S()
            opt = torch.optim.Adam(model.parameters(), lr=LR)
            loss, preds = script_loss(name, model, d, F)
            opt.zero_grad()
            loss.backward()
            grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
            opt.step()
            out[f"{sname}_loss"] = np.array(loss.item(), dtype=np.float64)
            for k, pr in enumerate(preds):
                out[f"{sname}_pred{k}"] = pr.detach().numpy().copy()
            for k, gv in grads.items():
                out[f"{sname}_g_{k}"] = gv
            for k, v in model.state_dict().items():
                out[f"{sname}_after_{k}"] = v.detach().numpy().copy()
            # fp64 restatement through the same reference code (the arbiter for gradient tolerances)
            m64 = copy.deepcopy(base).double()
            m64.solver = S()
            d64 = {k: v.double() for k, v in d.items()}
            loss64, preds64 = script_loss(name, m64, d64, F)
            loss64.backward()
            out[f"{sname}_loss64"] = np.array(loss64.item(), dtype=np.float64)
            for k, pr in enumerate(preds64):
                out[f"{sname}_pred64_{k}"] = pr.detach().numpy().copy()
            for k, p in m64.named_parameters():
                if p.grad is not None:
                    out[f"{sname}_g64_{k}"] = p.grad.detach().numpy().copy()
        path = os.path.join(HERE, f"script_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
