"""ctypes binding of libpsnode_b200.so (the C ABI in include/psnode_b200.h).

The structures below mirror the header field for field.  Loading is lazy and LOUD: if the shared library is
missing or cannot be loaded, `lib()` raises -- there is no Python / CPU fallback for the integration path.
"""
import ctypes as C
import os
import threading

PSNODE_MAX_LAYERS = 8
ABI_VERSION = 5

OK, EINVAL, EUNSUPPORTED, EWORKSPACE, ECUDA, ENODEVICE = 0, -1, -2, -3, -4, -5
EULER, MIDPOINT, RK4 = 0, 1, 2
ODE, DAE = 0, 1
IMPL_AUTO, IMPL_GENERIC, IMPL_FUSED, IMPL_TC, IMPL_TC8, IMPL_WIDE, IMPL_LAYER = 0, 1, 2, 3, 4, 5, 6
IMPL_BY_NAME = {"auto": IMPL_AUTO, "generic": IMPL_GENERIC, "fused": IMPL_FUSED, "tc": IMPL_TC, "tc8": IMPL_TC8, "wide": IMPL_WIDE, "layer": IMPL_LAYER}

_fp = C.POINTER(C.c_float)


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32),
                ("in_dim", C.c_int32 * PSNODE_MAX_LAYERS),
                ("out_dim", C.c_int32 * PSNODE_MAX_LAYERS),
                ("W", C.c_void_p * PSNODE_MAX_LAYERS),
                ("b", C.c_void_p * PSNODE_MAX_LAYERS)]


class Series(C.Structure):
    _fields_ = [("p", C.c_void_p), ("st", C.c_int64), ("sb", C.c_int64)]


class Problem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("method", C.c_int32), ("impl", C.c_int32),
                ("B", C.c_int32), ("T", C.c_int32),
                ("X", C.c_int32), ("Z", C.c_int32), ("V", C.c_int32), ("I", C.c_int32),
                ("teacher_x", C.c_int32), ("teacher_i", C.c_int32), ("E", C.c_int32),
                ("t", Series), ("x", Series), ("z", Series), ("v", Series), ("i", Series),
                ("x_init", C.c_void_p), ("x_init_sb", C.c_int64),
                ("a0", C.c_void_p), ("a0_sb", C.c_int64),
                ("event_idx", C.c_void_p),
                ("z_jump", C.c_void_p), ("zj_sb", C.c_int64), ("zj_se", C.c_int64),
                ("v_jump", C.c_void_p), ("vj_sb", C.c_int64), ("vj_se", C.c_int64),
                ("de", Mlp), ("ae", Mlp),
                ("x_sol", Series), ("i_sol", Series),
                ("tape", C.c_void_p), ("tape_floats", C.c_int64)]


class LossTerm(C.Structure):
    _fields_ = [("target", Series), ("mask", Series), ("feat_weight", C.c_void_p), ("scale", C.c_void_p)]


class Adjoint(C.Structure):
    _fields_ = [("gx", Series), ("gi", Series),
                ("d_theta", C.c_void_p), ("n_theta", C.c_int64),
                ("d_x0", C.c_void_p), ("d_x0_sb", C.c_int64),
                ("d_a0", C.c_void_p), ("d_a0_sb", C.c_int64),
                ("d_z", Series), ("d_v", Series),
                ("d_zjump", C.c_void_p), ("d_zj_sb", C.c_int64), ("d_zj_se", C.c_int64),
                ("d_vjump", C.c_void_p), ("d_vj_sb", C.c_int64), ("d_vj_se", C.c_int64),
                ("d_xteach", Series), ("d_iteach", Series),
                ("fuse_x", LossTerm), ("fuse_i", LossTerm)]


class Codec(C.Structure):
    _fields_ = [("ZR", C.c_int32), ("VR", C.c_int32), ("XR", C.c_int32), ("IR", C.c_int32),
                ("z_raw", Series), ("v_raw", Series),
                ("zj_raw", C.c_void_p), ("zjr_sb", C.c_int64), ("zjr_se", C.c_int64),
                ("vj_raw", C.c_void_p), ("vjr_sb", C.c_int64), ("vjr_se", C.c_int64),
                ("z_enc", Mlp), ("v_enc", Mlp), ("x_dec", Mlp), ("i_dec", Mlp),
                ("x_out", Series), ("i_out", Series),
                ("chunk_rows", C.c_int32)]


# every symbol include/psnode_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("psnode_abi_version", C.c_int, []),
    ("psnode_status_string", C.c_char_p, [C.c_int]),
    ("psnode_last_cuda_error", C.c_char_p, []),
    ("psnode_launch_count", C.c_int64, []),
    ("psnode_last_kernel", C.c_char_p, []),
    ("psnode_mlp_param_count", C.c_int64, [C.POINTER(Mlp)]),
    ("psnode_event_table", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    ("psnode_tape_floats", C.c_int64, [C.POINTER(Problem)]),
    ("psnode_tape_covers_input_grads", C.c_int, [C.POINTER(Problem)]),
    ("psnode_sweep_fuses_loss", C.c_int, [C.POINTER(Problem), C.POINTER(Adjoint)]),
    ("psnode_forward_workspace", C.c_int64, [C.POINTER(Problem)]),
    ("psnode_backward_workspace", C.c_int64, [C.POINTER(Problem), C.POINTER(Adjoint)]),
    ("psnode_forward", C.c_int, [C.POINTER(Problem), C.c_void_p, C.c_int64, C.c_void_p]),
    ("psnode_backward", C.c_int, [C.POINTER(Problem), C.POINTER(Adjoint), C.c_void_p, C.c_int64, C.c_void_p]),
    ("psnode_forward_host", C.c_int, [C.POINTER(Problem), C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("psnode_masked_sse_workspace", C.c_int64, []),
    ("psnode_masked_sse", C.c_int, [C.POINTER(Series), C.POINTER(Series), C.POINTER(Series), C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("psnode_masked_sse_grad", C.c_int, [C.POINTER(Series), C.POINTER(Series), C.POINTER(Series), C.c_void_p, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_void_p, C.POINTER(Series), C.c_void_p]),
    ("psnode_forward_encoded_workspace", C.c_int64, [C.POINTER(Problem), C.POINTER(Codec)]),
    ("psnode_forward_encoded", C.c_int, [C.POINTER(Problem), C.POINTER(Codec), C.c_void_p, C.c_int64, C.c_void_p]),
    ("psnode_init_state", C.c_int, [C.POINTER(Mlp), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    ("psnode_init_state_backward_workspace", C.c_int64, [C.POINTER(Mlp), C.c_int32]),
    ("psnode_init_state_backward", C.c_int, [C.POINTER(Mlp), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
]

# PSNODE_B200_LIB selects an alternative build of the SAME library (A/B kernel experiments); never a different backend.
LIB_PATH = os.environ.get("PSNODE_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libpsnode_b200.so")

_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise NativeLibraryError if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  py_psnode_b200 has no CPU fallback for the integration path.")
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as exc:
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
        for name, restype, argtypes in SYMBOLS:
            try:
                fn = getattr(handle, name)
            except AttributeError as exc:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.psnode_abi_version() != ABI_VERSION:
            raise NativeLibraryError(f"ABI mismatch: library {handle.psnode_abi_version()}, binding {ABI_VERSION}")
        _lib = handle
    return _lib


def check(status, what):
    if status == OK:
        return
    L = lib()
    msg = L.psnode_status_string(status).decode()
    if status == ECUDA:
        msg += ": " + L.psnode_last_cuda_error().decode()
    raise RuntimeError(f"{what} failed: {msg} (status {status})")


def launch_count():
    return int(lib().psnode_launch_count())


def last_kernel():
    return lib().psnode_last_kernel().decode()
