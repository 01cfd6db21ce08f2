"""py_psnode_b200 -- B200-native fixed-grid neural ODE/DAE integrator behind Py_PSNODE's `neural_dae` call surface.

    from py_psnode_b200 import RK4            # or: from neural_dae import RK4  (drop-in shim at the repo root)
    x_sol = RK4().integrate_ODE(x_func=de_func, t=t, x=x, z=z, all_initial=a0, event_fn=ev.event_fn,
                                jump_change_fn=ev.jump_change_fn)

The time loop, the RK stages, the small ELU-MLP right-hand side, event jumps and the trajectory write-back run in one
persistent CUDA kernel per call (csrc/, built into _lib/libpsnode_b200.so by `__graft_entry__.build()`).
"""
from .solvers import FixedGridODESolver, Euler, Midpoint, RK4
from .neural_base import (ODE_Curves_Sample, ODE_Event, DE_Func, ODE_Base,
                          DAE_Curves_Sample, DAE_Event, AE_Func, DAE_Base)
from .pattern import UnsupportedModuleError
from .utils import Logger, Losses

__all__ = ["FixedGridODESolver", "Euler", "Midpoint", "RK4", "ODE_Curves_Sample", "ODE_Event", "DE_Func", "ODE_Base",
           "DAE_Curves_Sample", "DAE_Event", "AE_Func", "DAE_Base", "UnsupportedModuleError", "Logger", "Losses"]
__version__ = "0.1.0"
