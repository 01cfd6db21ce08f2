"""Drop-in replacements for the reference's fixed-grid solver objects.

Same construction, attributes and call signatures as `neural_dae.my_solvers.FixedGridODESolver` and
`neural_dae.my_fixed_grid.{Euler, Midpoint, RK4}` (reference: neural_dae/my_solvers.py:8-131,
neural_dae/my_fixed_grid.py:12-59), but `integrate_ODE` / `integrate_DAE` run the whole time loop in ONE
persistent CUDA kernel (libpsnode_b200.so) instead of a Python loop that launches ~70 kernels per step.

What is kept from the reference, deliberately:
  * constructor kwargs `step_size`, `grid_constructor`, `interp` are accepted and inert (the reference never calls
    the grid constructor: my_solvers.py:54, :86 are commented out); passing both raises ValueError (:28-29);
  * the solution has the shape of `x` (`x_init` width when x is zero-width, :97), a fresh tensor on x.device;
  * RK4 is the 3/8-rule variant (my_fixed_grid.py:38-51), stage inputs z/v/i are zero-order held (:66, :104);
  * the event predicate looks at sample 0 only, with exact float equality (neural_base.py:54).
What differs: x_func / i_func must be Linear/ELU chains of the reference's DE_Func / AE_Func shape and the event
callbacks must be bound methods of an ODE_Event / DAE_Event (see pattern.py); anything else raises
UnsupportedModuleError unless the solver was built with `eager=True`, which runs the plain PyTorch loop.
"""
import abc
from typing import Optional

import torch
import torch.nn as nn

from . import _native as N
from . import engine, pattern

_ONE_THIRD = 1 / 3


def _params(layers):
    out = []
    for lin in layers:
        out.extend((lin.weight, lin.bias))
    return out


class FixedGridODESolver(metaclass=abc.ABCMeta):
    order: int
    _method: int

    def __init__(self, step_size=None, grid_constructor=None, interp="linear", impl: str = "auto", eager: bool = False,
                 check_events: bool = False):
        if step_size is not None and grid_constructor is not None:
            raise ValueError("step_size and grid_constructor are mutually exclusive arguments.")
        self.step_size = step_size
        self.interp = interp
        self.grid_constructor = grid_constructor if grid_constructor is not None else (lambda func, x0, t: t)
        # timing fields the reference declares and never updates (my_solvers.py:15-18); kept for attribute parity
        self.enable_cal_time = False
        self.assert_time = 0
        self.cal_time = 0
        self.total_time = 0
        if impl not in N.IMPL_BY_NAME:
            raise ValueError(f"impl must be one of {sorted(N.IMPL_BY_NAME)}")
        self.impl = impl
        self.eager = eager
        self.check_events = check_events

    # ------------------------------------------------------------------ single step (public in the reference)
    @abc.abstractmethod
    def _step_func(self, func, t0, dt, t1, x0, z0=None, v0=None, i0=None, all_initial=None):
        """Return (dx, f0) for one step; evaluates `func` with the reference's keyword contract."""

    @staticmethod
    def _rhs(func, t0, x, z0, v0, i0, all_initial):
        if v0 is None:
            return func(t0=t0, xt=x, zt=z0, all_initial=all_initial)
        return func(t0=t0, xt=x, zt=z0, vt=v0, it=i0, all_initial=all_initial)

    def step_integrate(self, func, t0, dt, t1, x0, z0=None, v0=None, i0=None, all_initial=None):
        """(x0 + dx, f0): one explicit step with an arbitrary module (my_solvers.py:48-50)."""
        dx, f0 = self._step_func(func=func, t0=t0, dt=dt, t1=t1, x0=x0, z0=z0, v0=v0, i0=i0, all_initial=all_initial)
        return x0 + dx, f0

    # ------------------------------------------------------------------ fused integration
    def integrate_ODE(self, x_func: nn.Module, t: torch.Tensor, x: torch.Tensor, z: torch.Tensor, all_initial: torch.Tensor,
                      event_fn=None, jump_change_fn=None, input_true_x=False):
        """x_solution (T,B,X) for dx/dt = x_func(x, z held, all_initial) on the fixed grid t (my_solvers.py:52-80)."""
        if self.eager:
            return self._eager_ode(x_func, t, x, z, all_initial, event_fn, jump_change_fn, input_true_x)
        X, Z = x.shape[-1], z.shape[-1]
        de = pattern.match_de(x_func, X=X, Z=Z, dae=False)
        ev = pattern.match_event(event_fn, jump_change_fn, dae=False)
        cfg = engine.Config(kind=N.ODE, method=self._method, impl=N.IMPL_BY_NAME[self.impl], X=X, Z=Z, V=0, I=0,
                            teacher_x=bool(input_true_x), teacher_i=False, n_de=len(de), n_ae=0,
                            has_event=ev is not None, check_events=self.check_events)
        tens = [t, x, z, None, None, None, all_initial,
                ev[0] if ev else None, ev[1] if ev else None, None, *_params(de)]
        x_sol, _ = engine.integrate(cfg, tens)
        return x_sol

    def integrate_DAE(self, x_init: torch.Tensor, x_func: nn.Module, i_func: nn.Module, t: torch.Tensor, x: torch.Tensor,
                      z: torch.Tensor, v: torch.Tensor, i: torch.Tensor, all_initial: torch.Tensor, event_fn=None,
                      jump_change_fn=None, input_true_x=False, input_true_i=False):
        """(x_solution (T,B,X), i_solution (T,B,I)); i = i_func(x, z, v) evaluated explicitly once per step
        (my_solvers.py:82-131 -- the reference has no Newton iteration)."""
        if self.eager:
            return self._eager_dae(x_init, x_func, i_func, t, x, z, v, i, all_initial, event_fn, jump_change_fn,
                                   input_true_x, input_true_i)
        X, Z, V, I = x_init.shape[-1], z.shape[-1], v.shape[-1], i.shape[-1]
        if input_true_x and x.shape[-1] != X:
            raise ValueError("input_true_x needs a ground-truth x series of the state width")
        de = pattern.match_de(x_func, X=X, Z=Z, V=V, I=I, dae=True)
        ae = pattern.match_ae(i_func, X=X, Z=Z, V=V, I=I)
        ev = pattern.match_event(event_fn, jump_change_fn, dae=True)
        cfg = engine.Config(kind=N.DAE, method=self._method, impl=N.IMPL_BY_NAME[self.impl], X=X, Z=Z, V=V, I=I,
                            teacher_x=bool(input_true_x), teacher_i=bool(input_true_i), n_de=len(de), n_ae=len(ae),
                            has_event=ev is not None, check_events=self.check_events)
        tens = [t, x if x.shape[-1] != 0 else None, z, v, i, x_init, all_initial,
                ev[0] if ev else None, ev[1] if ev else None, ev[2] if ev else None, *_params(de), *_params(ae)]
        return engine.integrate(cfg, tens)

    # ------------------------------------------------------------------ opt-in eager loop (never chosen automatically)
    def _eager_ode(self, x_func, t, x, z, all_initial, event_fn, jump_change_fn, input_true_x):
        rows = [x[0]]
        prev = x[0]
        for j in range(1, t.shape[0]):
            t0, t1, z0 = t[j - 1], t[j], z[j - 1]
            if event_fn is not None and event_fn(t0) == True:   # noqa: E712  (callbacks may return tensors)
                z0 = jump_change_fn(t0, z0)
            start = x[j - 1] if input_true_x else prev
            prev, _ = self.step_integrate(func=x_func, t0=t0, dt=t1 - t0, t1=t1, x0=start, z0=z0, all_initial=all_initial)
            rows.append(prev)
        return torch.stack(rows, dim=0)

    def _eager_dae(self, x_init, x_func, i_func, t, x, z, v, i, all_initial, event_fn, jump_change_fn, input_true_x,
                   input_true_i):
        x_prev = x_init
        i_prev = i_func(xt=x[0] if input_true_x else x_prev, zt=z[0], vt=v[0], all_initial=all_initial)
        xr, ir = [x_prev], [i_prev]
        for j in range(1, t.shape[0]):
            t0, t1, z0, v0 = t[j - 1], t[j], z[j - 1], v[j - 1]
            if event_fn is not None and event_fn(t0) == True:   # noqa: E712
                z0, v0 = jump_change_fn(t0, z0, v0)
                i_prev = i_func(xt=x_prev, zt=z0, vt=v0, all_initial=all_initial)
            start = x[j - 1] if input_true_x else x_prev
            held = i[j - 1] if input_true_i else i_prev
            x_prev, _ = self.step_integrate(func=x_func, t0=t0, dt=t1 - t0, t1=t1, x0=start, z0=z0, v0=v0, i0=held,
                                            all_initial=all_initial)
            i_prev = i_func(xt=x[j] if input_true_x else x_prev, zt=z[j], vt=v[j], all_initial=all_initial)
            xr.append(x_prev)
            ir.append(i_prev)
        return torch.stack(xr, dim=0), torch.stack(ir, dim=0)


class Euler(FixedGridODESolver):
    order = 1
    _method = N.EULER

    def _step_func(self, func, t0, dt, t1, x0, z0=None, v0=None, i0=None, all_initial=None):
        f0 = self._rhs(func, t0, x0, z0, v0, i0, all_initial)
        return dt * f0, f0


class Midpoint(FixedGridODESolver):
    order = 2
    _method = N.MIDPOINT

    def _step_func(self, func, t0, dt, t1, x0, z0=None, v0=None, i0=None, all_initial=None):
        half_dt = 0.5 * dt
        f0 = self._rhs(func, t0, x0, z0, v0, i0, all_initial)
        f_mid = self._rhs(func, t0 + half_dt, x0 + f0 * half_dt, z0, v0, i0, all_initial)
        return dt * f_mid, f0


class RK4(FixedGridODESolver):
    """Fourth order, 3/8 rule (the reference's `rk4_alt_step_func`)."""
    order = 4
    _method = N.RK4

    def _step_func(self, func, t0, dt, t1, x0, z0=None, v0=None, i0=None, all_initial=None):
        k1 = self._rhs(func, t0, x0, z0, v0, i0, all_initial)
        k2 = self._rhs(func, t0 + dt * _ONE_THIRD, x0 + dt * k1 * _ONE_THIRD, z0, v0, i0, all_initial)
        k3 = self._rhs(func, t0 + dt * (2 * _ONE_THIRD), x0 + dt * (k2 - k1 * _ONE_THIRD), z0, v0, i0, all_initial)
        k4 = self._rhs(func, t1, x0 + dt * (k1 - k2 + k3), z0, v0, i0, all_initial)
        return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125, k1
