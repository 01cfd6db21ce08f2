"""`neural_dae.neural_base` (imported as a submodule by neural_01_DAE_01_no_encode.py:5)."""
from py_psnode_b200.neural_base import *                                                    # noqa: F401,F403
from py_psnode_b200.neural_base import (ODE_Curves_Sample, ODE_Event, DE_Func, ODE_Base,   # noqa: F401
                                        DAE_Curves_Sample, DAE_Event, AE_Func, DAE_Base)
