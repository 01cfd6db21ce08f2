"""Support classes the reference's scripts import from `neural_dae` / `neural_dae.neural_base`
(reference: neural_dae/__init__.py:1-2, neural_dae/neural_base.py): event holders, `.npz` datasets and the
generic model shells.  Host-side code; the integration itself lives in solvers.py / libpsnode_b200.so.
"""
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import Dataset

from .solvers import Euler, FixedGridODESolver


# ----------------------------------------------------------------------------------------------- events
class _EventBase:
    """Step-jump events of the external inputs (reference: neural_base.py:43-65, 169-196).

    `event_t` is (B, E, 1); event k of EVERY sample fires when the grid time of SAMPLE 0 equals event_t[0, k]
    exactly.  The fused integrator never calls these methods: it reads `event_t` / `*_jump` off the object
    (pattern.match_event) and builds a per-step table on the device.  They remain callable for user code (the
    reference's public callback contract), nothing on the integration path uses them.
    """
    _jump_names = ()

    def __init__(self):
        self.event_t: Optional[torch.Tensor] = None
        for n in self._jump_names:
            setattr(self, n, None)

    def _match(self, t0: torch.Tensor) -> torch.Tensor:
        return (self.event_t[0].reshape(-1) == t0.reshape(-1)[0]).nonzero().reshape(-1)

    def event_fn(self, t0: torch.Tensor) -> bool:
        if self.event_t is None:
            return False
        return bool(self._match(t0).numel() > 0)

    def _jumped(self, t0: torch.Tensor, held: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
        hit = self._match(t0)
        if hit.numel() != 1:
            raise RuntimeError(f"{hit.numel()} events match grid time {float(t0.reshape(-1)[0])}; exactly one is required")
        return table[:, int(hit[0])].reshape(held.shape)


class ODE_Event(_EventBase):
    _jump_names = ("z_jump",)

    def set_event(self, t: torch.Tensor, z: torch.Tensor):
        self.event_t, self.z_jump = t, z

    def jump_change_fn(self, t0: torch.Tensor, z0: torch.Tensor):
        return self._jumped(t0, z0, self.z_jump)


class DAE_Event(_EventBase):
    _jump_names = ("z_jump", "v_jump")

    def set_event(self, t: torch.Tensor, z: torch.Tensor, v: torch.Tensor):
        self.event_t, self.z_jump, self.v_jump = t, z, v

    def jump_change_fn(self, t0, z0, v0):
        return self._jumped(t0, z0, self.z_jump), self._jumped(t0, v0, self.v_jump)


# ----------------------------------------------------------------------------------------------- datasets
class _CurveSamples(Dataset):
    """`.npz` trajectory sets (reference: neural_base.py:10-40, 136-166).  Keys: t, x, z[, v, i], event_t, z_jump
    [, v_jump], mask (optional for ODE sets), name.  Sub-sampling uses default_rng(42) like the reference so the
    same rows are drawn; `cut_length` truncates the time axis; `contain_larger_than` redraws until some x exceeds it.
    """
    series = ()
    per_sample = ()
    mask_optional = False

    def __init__(self, data_path, device=None, num_sample=None, cut_length=None, contain_larger_than=None):
        super().__init__()
        self.rng = np.random.default_rng(42)
        f = np.load(data_path, allow_pickle=True)
        total = f["t"].shape[0]
        while True:
            index = np.arange(total)
            if num_sample is not None:
                index = self.rng.choice(index, num_sample, replace=False)
            if contain_larger_than is None or np.any(f["x"][index] > contain_larger_than):
                break
        steps = f["t"].shape[1] if cut_length is None else min(cut_length, f["t"].shape[1])
        self.data_name = f["name"] if "name" in f.files else None
        for key in self.series:
            setattr(self, key, torch.from_numpy(f[key][index][:, :steps]))
        for key in self.per_sample:
            setattr(self, key, torch.from_numpy(f[key][index]))
        if "mask" in f.files:
            self.mask = torch.from_numpy(f["mask"][index][:, :steps])
        elif self.mask_optional:
            self.mask = torch.ones(self.x.shape, dtype=torch.float32)
        else:
            raise KeyError("mask")
        lengths = {getattr(self, k).shape[1] for k in self.series}
        assert len(lengths) == 1, "Sample shapes are wrong!"

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, idx):
        return tuple(getattr(self, k)[idx] for k in (*self.series, *self.per_sample, "mask"))


class ODE_Curves_Sample(_CurveSamples):
    series = ("t", "x", "z")
    per_sample = ("event_t", "z_jump")
    mask_optional = True


class DAE_Curves_Sample(_CurveSamples):
    series = ("t", "x", "z", "v", "i")
    per_sample = ("event_t", "z_jump", "v_jump")


# ----------------------------------------------------------------------------------------------- model shells
def _elu_mlp(widths):
    mods = []
    for k in range(len(widths) - 1):
        mods.append(nn.Linear(widths[k], widths[k + 1]))
        if k != len(widths) - 2:
            mods.append(nn.ELU())
    return nn.Sequential(*mods)


class DE_Func(nn.Module):
    """dx/dt network on cat(a0, s - a0, s), s = cat(x, z[, v, i]).  `depth` Linear layers (4 = *_01 scripts, 2 = *_02).
    The same-named class in the reference's neural_base.py (:68-115) is stale dead code; this one
    follows the live script-local classes so the name stays importable AND usable."""

    def __init__(self, x_dim: int, z_dim: int, hidden_dim: int, v_dim: int = 0, i_dim: int = 0, depth: int = 4):
        super().__init__()
        s = x_dim + z_dim + v_dim + i_dim
        self.x_dot = _elu_mlp([3 * s] + [hidden_dim] * (depth - 1) + [x_dim])

    def forward(self, t0, xt, zt, all_initial, vt=None, it=None):
        parts = (xt, zt) if vt is None else (xt, zt, vt, it)
        s = torch.cat(parts, dim=-1)
        return self.x_dot(torch.cat((all_initial, s - all_initial, s), dim=-1))


class AE_Func(nn.Module):
    """Algebraic network i = mlp(cat(a0, x, z, v)) (live form: neural_01_DAE_01_no_encode.py:74-83)."""

    def __init__(self, x_dim: int, v_dim: int, i_dim: int, hidden_dim: int, z_dim: int = 0, depth: int = 4):
        super().__init__()
        s = x_dim + z_dim + v_dim + i_dim
        self.i_calculator = _elu_mlp([s + x_dim + z_dim + v_dim] + [hidden_dim] * (depth - 1) + [i_dim])

    def forward(self, xt, zt, vt, all_initial):
        return self.i_calculator(torch.cat((all_initial, xt, zt, vt), dim=-1))


class ODE_Base(nn.Module):
    """Minimal ODE model shell: batch-major (B,T,.) in, batch-major prediction out."""

    def __init__(self, x_dim: int, z_dim: int, hidden_dim: int, solver: Optional[FixedGridODESolver] = None):
        super().__init__()
        self.de_func = DE_Func(x_dim=x_dim, z_dim=z_dim, hidden_dim=hidden_dim)
        self.solver = solver if solver is not None else Euler()
        self.event = ODE_Event()

    def forward(self, t, x, z, event_t=None, z_jump=None):
        if event_t is not None:
            self.event.set_event(t=event_t, z=z_jump)
        xt, zt = x.permute(1, 0, 2), z.permute(1, 0, 2)
        a0 = torch.cat((xt[0], zt[0]), dim=-1)
        sol = self.solver.integrate_ODE(x_func=self.de_func, t=t.permute(1, 0, 2), x=xt, z=zt, all_initial=a0,
                                        event_fn=self.event.event_fn, jump_change_fn=self.event.jump_change_fn)
        return sol.permute(1, 0, 2)


class DAE_Base(nn.Module):
    """Minimal DAE model shell (initial state taken from x[:,0])."""

    def __init__(self, x_dim: int, z_dim: int, v_dim: int, i_dim: int, hidden_dim: int,
                 solver: Optional[FixedGridODESolver] = None):
        super().__init__()
        self.de_func = DE_Func(x_dim=x_dim, z_dim=z_dim, hidden_dim=hidden_dim, v_dim=v_dim, i_dim=i_dim)
        self.ae_func = AE_Func(x_dim=x_dim, v_dim=v_dim, i_dim=i_dim, hidden_dim=hidden_dim, z_dim=z_dim)
        self.solver = solver if solver is not None else Euler()
        self.event = DAE_Event()

    def forward(self, t, x, z, v, i, event_t=None, z_jump=None, v_jump=None):
        if event_t is not None:
            self.event.set_event(t=event_t, z=z_jump, v=v_jump)
        xt, zt, vt, it = (q.permute(1, 0, 2) for q in (x, z, v, i))
        a0 = torch.cat((xt[0], zt[0], vt[0], it[0]), dim=-1)
        xs, is_ = self.solver.integrate_DAE(x_init=xt[0], x_func=self.de_func, i_func=self.ae_func, t=t.permute(1, 0, 2),
                                            x=xt, z=zt, v=vt, i=it, all_initial=a0, event_fn=self.event.event_fn,
                                            jump_change_fn=self.event.jump_change_fn)
        return xs.permute(1, 0, 2), is_.permute(1, 0, 2)
