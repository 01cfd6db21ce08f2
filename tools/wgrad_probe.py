"""Decode what psn_lg_wgrad_kernel computes (debugging aid for the MN-major descriptors)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
from py_psnode_b200 import _native
lib = _native.lib()
fn = lib.psnode_debug_lg_wgrad
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
wsf = lib.psnode_debug_lg_wgrad_workspace; wsf.restype = C.c_int64; wsf.argtypes = [C.c_int, C.c_int]
dev = "cuda:0"
M = K = 128; N = 32
def run(P, Q):
    out = torch.full((M, K), -7.0, device=dev)
    ws = torch.zeros(int(wsf(M, K)), dtype=torch.uint8, device=dev)
    st = fn(P.data_ptr(), P.stride(1), P.stride(0), M, Q.data_ptr(), Q.stride(1), Q.stride(0), K, 1, N, out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return st, out, ws
mode = os.environ.get("PSNODE_WG_DBG", "0")
torch.manual_seed(0)
P = torch.randn(1, N, M, device=dev); Q = torch.randn(1, N, K, device=dev)
if mode in ("1", "2"):
    src = P if mode == "1" else Q
    # encode (row, col) into the value to see the smem layout
    enc = (torch.arange(N, device=dev).view(N, 1) * 1000 + torch.arange(M, device=dev).view(1, M)).float().view(1, N, M)
    st, out, ws = run(enc if mode == "1" else P, enc if mode == "2" else Q)
    slab = ws[256:256 + 16384].view(torch.float32)
    print("status", st, "first 40 floats of the slab (value = 1000 * row + col):", slab[:40].tolist())
    print("float at byte 128 (row 1 start):", slab[32:36].tolist(), " at 1024 (row 8):", slab[256:260].tolist(), " at 4096 (feature block 1):", slab[1024:1028].tolist())
else:
    st, out, ws = run(P, Q)
    if mode == "3":
        print("dbg ints", ws[:16].view(torch.int32).tolist(), "floats", ws[16:48].view(torch.float32).tolist(), "P[0,0,0]", float(P[0,0,0]), "Q[0,0,0]", float(Q[0,0,0]))
        print("slab[0:4]", ws[256:272].view(torch.float32).tolist(), "out[0,:4]", out[0,:4].tolist())
    want = torch.einsum("snm,snk->mk", P.double(), Q.double())
    print("LBO", os.environ.get("PSNODE_WG_LBO"), "SBO", os.environ.get("PSNODE_WG_SBO"), "KADV", os.environ.get("PSNODE_WG_KADV"),
          ": max err", float((out.double() - want).abs().max()), "err vs transposed", float((out.double().T - want).abs().max()),
          "out abs mean", float(out.abs().mean()), "want abs mean", float(want.abs().mean()))
