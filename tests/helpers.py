"""Shared test utilities: golden fixture loading and module reconstruction."""
import glob
import os

import numpy as np
import torch
import torch.nn as nn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 1e-5, 1e-6          # BASELINE.json north_star: rtol=1e-5 / atol=1e-6 fp32


def golden_names(kind=None):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if kind is None:
        return names
    return [n for n in names if n.startswith(kind)]


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def params_of(d, prefix, dtype=torch.float32, device="cpu"):
    out, k = [], 0
    while f"{prefix}_W{k}" in d:
        out.append((torch.from_numpy(d[f"{prefix}_W{k}"]).to(device=device, dtype=dtype),
                    torch.from_numpy(d[f"{prefix}_b{k}"]).to(device=device, dtype=dtype)))
        k += 1
    return out


def _seq(params):
    mods = []
    for k, (W, b) in enumerate(params):
        lin = nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            lin.weight.copy_(W)
            lin.bias.copy_(b)
        mods.append(lin)
        if k != len(params) - 1:
            mods.append(nn.ELU())
    return nn.Sequential(*mods)


class DE(nn.Module):
    """Module with the reference's DE_Func contract, rebuilt from stored weights."""

    def __init__(self, params):
        super().__init__()
        self.x_dot = _seq(params)

    def forward(self, t0, xt, zt, all_initial, vt=None, it=None):
        s = torch.cat((xt, zt) if vt is None else (xt, zt, vt, it), dim=-1)
        return self.x_dot(torch.cat((all_initial, s - all_initial, s), dim=-1))


class AE(nn.Module):
    def __init__(self, params):
        super().__init__()
        self.i_calculator = _seq(params)

    def forward(self, xt, zt, vt, all_initial):
        return self.i_calculator(torch.cat((all_initial, xt, zt, vt), dim=-1))


def tm(a, device="cpu", dtype=torch.float32):
    """batch-major numpy (B,T,W) -> time-major *view* (T,B,W), like the reference's x.permute(1,0,2)."""
    return torch.from_numpy(a).to(device=device, dtype=dtype).permute(1, 0, 2)


def tol_report(got, want, want64=None):
    diff = (got.double() - want.double()).abs()
    msg = f"max|got-ref32|={diff.max().item():.3e}"
    if want64 is not None:
        msg += (f" max|got-ref64|={(got.double() - want64).abs().max().item():.3e}"
                f" max|ref32-ref64|={(want.double() - want64).abs().max().item():.3e}")
    return msg
