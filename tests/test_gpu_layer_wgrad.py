"""The MN-major tcgen05 weight-gradient GEMM of the layer path (psn_lg_wgrad_kernel, csrc/psnode_lg.cu) alone, through a test
hook of the shared library: out = sum over slots and rows of P^T Q with both operands in their (row, feature) layout, against
float64 einsum.  Covers row counts that are not multiples of the 32-row chunk, several slots and strided slot views."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,nslots,N", [(128, 128, 1, 64), (256, 256, 3, 200), (256, 256, 16, 1000), (128, 256, 2, 37)])
def test_layer_wgrad_gemm_against_einsum(native_lib, M, K, nslots, N):
    from py_psnode_b200 import _native
    lib = _native.lib()
    fn = lib.psnode_debug_lg_wgrad
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                   C.c_void_p, C.c_int64, C.c_void_p]
    wsf = lib.psnode_debug_lg_wgrad_workspace
    wsf.restype = C.c_int64
    wsf.argtypes = [C.c_int, C.c_int]
    torch.manual_seed(3)
    dev = "cuda:0"
    P = torch.randn(nslots, N, M, device=dev) * 0.3
    Qfull = torch.randn(nslots, N, K + 64, device=dev) * 0.3          # Q is a strided view (row stride K + 64)
    Q = Qfull[:, :, :K]
    out = torch.empty(M, K, device=dev)
    ws = torch.empty(int(wsf(M, K)), dtype=torch.uint8, device=dev)
    st = fn(P.data_ptr(), P.stride(1), P.stride(0), M, Q.data_ptr(), Q.stride(1), Q.stride(0), K, nslots, N, out.data_ptr(), ws.data_ptr(),
            ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert st == 0, st
    torch.cuda.synchronize()
    want = torch.einsum("snm,snk->mk", P.double(), Q.double())
    err = float((out.double() - want).abs().max())
    scale = float(want.abs().max())
    assert err <= 2e-6 * scale, f"max err {err:.3e} vs scale {scale:.3e}"
