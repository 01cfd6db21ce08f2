"""`neural_dae.my_fixed_grid`: the three fixed-step schemes."""
from py_psnode_b200.solvers import Euler, Midpoint, RK4                                     # noqa: F401
