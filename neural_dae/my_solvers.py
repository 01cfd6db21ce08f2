"""`neural_dae.my_solvers`: the solver base class."""
from py_psnode_b200.solvers import FixedGridODESolver                                       # noqa: F401
