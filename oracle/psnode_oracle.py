"""CPU oracle for the fixed-grid neural ODE/DAE integration path.  TEST INFRASTRUCTURE ONLY.

This is a from-scratch restatement, in plain PyTorch-CPU tensor ops, of what the reference
(xxh0523/Py_PSNODE @ d366e75) computes on its hot path.  It exists so that the CUDA path can be
checked and so that `bench.py` has a CPU baseline on a box where /root/reference is absent.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it; nothing under `py_psnode_b200/` or `neural_dae/` does (a test enforces that).

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself: `tests/golden/*.npz`, produced by
`tests/golden/make_golden.py` running the unmodified reference in the build container.
`tests/test_oracle_golden.py` requires this file to reproduce every one of them (forward fp32,
forward fp64 and all autograd gradients).

Why torch ops and not numpy/C: the reference's arithmetic IS ATen-CPU (`addmm`, vectorised
`elu` with expm1 semantics; torch is the un-vendored third-party dependency,
version 2.11.0+cu128 in this image, the reference pins none).  Using the same library kernels in
the same order makes the fp32 oracle reproduce the reference bit-for-bit on the same host, and
makes the timed CPU baseline the same work the reference's loop does.  A float64 run of the same
functions (pass `.double()` tensors) is the tie-breaker between two fp32 results.

Reference lines restated (all paths relative to /root/reference):
  mlp()                  nn.Sequential(Linear, ELU, ..., Linear)   neural_00_ODE_01_no_encode.py:61-64 (4 Linear),
                                                                   neural_00_ODE_02_direct_encode.py:52-53 (2 Linear)
  de_rhs()               DE_Func.forward                           neural_00_ODE_01_no_encode.py:66-68,
                                                                   neural_01_DAE_01_no_encode.py:69-71
  ae_eval()              AE_Func.forward                           neural_01_DAE_01_no_encode.py:82-83
  increment()            Euler/Midpoint/RK4._step_func             neural_dae/my_fixed_grid.py:15-18, 23-32, 38-59
  event_index()          ODE_Event/DAE_Event.event_fn              neural_dae/neural_base.py:52-57, 180-185
  (jump selection)       ODE_Event/DAE_Event.jump_change_fn        neural_dae/neural_base.py:59-65, 187-196
  integrate_ode()        FixedGridODESolver.integrate_ODE          neural_dae/my_solvers.py:52-80
  integrate_dae()        FixedGridODESolver.integrate_DAE          neural_dae/my_solvers.py:82-131
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Sequence[Tuple[torch.Tensor, torch.Tensor]]   # [(W[out,in], b[out]), ...] in layer order

ONE_THIRD = 1 / 3     # Python doubles, multiplied into tensors exactly like my_fixed_grid.py:8-9


def mlp(params: Params, u: torch.Tensor) -> torch.Tensor:
    """Linear -> ELU -> ... -> Linear (no activation after the last layer)."""
    last = len(params) - 1
    for li, (W, b) in enumerate(params):
        u = F.linear(u, W, b)
        if li != last:
            u = F.elu(u)
    return u


def de_rhs(de: Params, a0: torch.Tensor, x: torch.Tensor, held: Sequence[torch.Tensor]) -> torch.Tensor:
    """dx/dt network: input cat(a0, s - a0, s) with s = cat(x, held inputs)."""
    s = torch.cat((x, *held), dim=-1)
    return mlp(de, torch.cat((a0, s - a0, s), dim=-1))


def ae_eval(ae: Params, a0: torch.Tensor, x: torch.Tensor, z: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Algebraic network: input cat(a0, x, z, v); evaluated explicitly, never iterated."""
    return mlp(ae, torch.cat((a0, x, z, v), dim=-1))


def increment(method: str, f, x0: torch.Tensor, dt: torch.Tensor) -> torch.Tensor:
    """One-step increment dx of the named scheme; `f(x)` is the RHS with everything else held."""
    if method == "euler":
        return dt * f(x0)
    if method == "midpoint":
        half_dt = 0.5 * dt
        f0 = f(x0)
        x_mid = x0 + f0 * half_dt
        return dt * f(x_mid)
    if method == "rk4":     # the 3/8-rule variant, operation order as in rk4_alt_step_func
        k1 = f(x0)
        k2 = f(x0 + dt * k1 * ONE_THIRD)
        k3 = f(x0 + dt * (k2 - k1 * ONE_THIRD))
        k4 = f(x0 + dt * (k1 - k2 + k3))
        return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
    raise ValueError(f"unknown method {method!r}")


def event_index(t_prev: torch.Tensor, event_t: Optional[torch.Tensor]) -> int:
    """Which event (if any) fires when leaving grid point t_prev (shape (B,1)).

    Only SAMPLE 0 is inspected and the comparison is exact float equality, as in the reference.
    Returns -1 for no event.  More than one match is an error (the reference's `.view` would throw).
    """
    if event_t is None:
        return -1
    hit = (event_t[0].reshape(-1) == t_prev[0][0]).nonzero().reshape(-1)
    if hit.numel() == 0:
        return -1
    if hit.numel() > 1:
        raise ValueError("more than one event matches the same grid time")
    return int(hit[0])


def integrate_ode(method: str, de: Params, t: torch.Tensor, x: torch.Tensor, z: torch.Tensor, a0: torch.Tensor,
                  event_t: Optional[torch.Tensor] = None, z_jump: Optional[torch.Tensor] = None,
                  teacher_x: bool = False) -> torch.Tensor:
    """t (T,B,1), x (T,B,X), z (T,B,Z) time-major views; a0 (B,X+Z); event_t (B,E,1), z_jump (B,E,Z)."""
    T = t.shape[0]
    rows: List[torch.Tensor] = [x[0]]
    prev = x[0]
    for j in range(1, T):
        dt = t[j] - t[j - 1]
        z0 = z[j - 1]
        k = event_index(t[j - 1], event_t)
        if k >= 0:
            z0 = z_jump[:, k, :]
        start = x[j - 1] if teacher_x else prev
        prev = start + increment(method, lambda xx: de_rhs(de, a0, xx, (z0,)), start, dt)
        rows.append(prev)
    return torch.stack(rows, dim=0)


def integrate_dae(method: str, de: Params, ae: Params, x_init: torch.Tensor, t: torch.Tensor, x: torch.Tensor,
                  z: torch.Tensor, v: torch.Tensor, i: torch.Tensor, a0: torch.Tensor,
                  event_t: Optional[torch.Tensor] = None, z_jump: Optional[torch.Tensor] = None,
                  v_jump: Optional[torch.Tensor] = None, teacher_x: bool = False, teacher_i: bool = False
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """x_init (B,X); series time-major; a0 (B,X+Z+V+I).  Returns (x_sol (T,B,X), i_sol (T,B,I))."""
    T = t.shape[0]
    x_prev = x_init
    i_prev = ae_eval(ae, a0, x[0] if teacher_x else x_prev, z[0], v[0])
    x_rows, i_rows = [x_prev], [i_prev]
    for j in range(1, T):
        dt = t[j] - t[j - 1]
        z0, v0 = z[j - 1], v[j - 1]
        k = event_index(t[j - 1], event_t)
        if k >= 0:
            z0, v0 = z_jump[:, k, :], v_jump[:, k, :]
            i_prev = ae_eval(ae, a0, x_prev, z0, v0)     # uses the PREDICTED state even under teacher forcing
        start = x[j - 1] if teacher_x else x_prev
        i_held = i[j - 1] if teacher_i else i_prev
        x_prev = start + increment(method, lambda xx: de_rhs(de, a0, xx, (z0, v0, i_held)), start, dt)
        i_prev = ae_eval(ae, a0, x[j] if teacher_x else x_prev, z[j], v[j])
        x_rows.append(x_prev)
        i_rows.append(i_prev)
    return torch.stack(x_rows, dim=0), torch.stack(i_rows, dim=0)


def params_from_npz(d, prefix: str, dtype=torch.float32) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """[(W,b)] from a golden fixture's `<prefix>_W<k>` / `<prefix>_b<k>` arrays."""
    out = []
    k = 0
    while f"{prefix}_W{k}" in d:
        out.append((torch.from_numpy(d[f"{prefix}_W{k}"]).to(dtype), torch.from_numpy(d[f"{prefix}_b{k}"]).to(dtype)))
        k += 1
    return out
