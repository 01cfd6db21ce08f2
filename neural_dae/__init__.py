"""Drop-in `neural_dae` package: the names Py_PSNODE's training scripts import, served by py_psnode_b200."""
from py_psnode_b200.neural_base import ODE_Curves_Sample, ODE_Event, DE_Func, ODE_Base      # noqa: F401
from py_psnode_b200.neural_base import DAE_Curves_Sample, DAE_Event, AE_Func, DAE_Base      # noqa: F401
from py_psnode_b200.solvers import Euler, Midpoint, RK4, FixedGridODESolver                # noqa: F401
from . import neural_base, my_solvers, my_fixed_grid                                       # noqa: F401
