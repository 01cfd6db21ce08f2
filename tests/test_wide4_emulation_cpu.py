"""CPU model of the tensor-core kernels of the 4-layer hidden-128 ODE_01 path (`tests/wide4_emulate.py`, test infrastructure like the oracle it checks against): the transliterated index
arithmetic of psn_wide4_fwd_kernel / psn_wide4_bwd_kernel -- operand tiles as byte arrays with the descriptor addressing of psnode_tc.cuh, TMEM
as a 128 x 512 array, the K-partial bookkeeping, the M = 64 row -> lane map, the slab assembly of the narrow gradients -- must reproduce the oracle's
trajectory and float64 autograd gradients.  This is how those kernels were written without a GPU at hand; it keeps the layout contract checkable
on the CPU (the GPU parity tests proper are tests/test_gpu_wide4.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_wide4_forward_index_arithmetic_reproduces_the_oracle():
    import wide4_emulate as E
    E.main(cases=(("rk4", 16, 2, 16, 2, 0, 128), ("midpoint", 5, 3, 19, 3, 1, 100)))


def test_wide4_reverse_sweep_index_arithmetic_reproduces_fp64_autograd():
    import wide4_emulate as E
    E.main_bwd(cases=(("rk4", 16, 2, 16, 2, 0, 128), ("euler", 5, 3, 7, 3, 1, 72)))
