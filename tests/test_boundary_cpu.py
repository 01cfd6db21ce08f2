"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/psnode_b200.h declares, the
ctypes structures match the header, the product never imports the oracle, and the product path fails LOUDLY without CUDA
(no CPU / Python-loop fallback).  No compute calls: there is no GPU here."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "psnode_b200.h")).read()


def test_library_exports_every_declared_symbol(native_lib):
    from py_psnode_b200 import _native
    text = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    declared = set(re.findall(r"\b(psnode_[a-z_0-9]+)\s*\(", text))
    assert declared, "no function declarations found in the header"
    bound = {name for name, _, _ in _native.SYMBOLS}
    assert declared == bound, f"header vs binding mismatch: {sorted(declared ^ bound)}"
    for name in declared:
        assert hasattr(native_lib, name), f"{name} not exported by libpsnode_b200.so"
    assert native_lib.psnode_abi_version() == _native.ABI_VERSION
    assert native_lib.psnode_status_string(-2).decode().startswith("problem shape not supported")
    assert native_lib.psnode_launch_count() >= 0


def test_ctypes_structs_match_header_sizes():
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    import subprocess
    import tempfile
    from py_psnode_b200 import _native
    src = ('#include <stdio.h>\n#include "psnode_b200.h"\nint main(void){printf("%zu %zu %zu %zu %d %d %zu\\n", sizeof(psnode_mlp), '
           'sizeof(psnode_series), sizeof(psnode_problem), sizeof(psnode_adjoint), PSNODE_IMPL_TC, PSNODE_MAX_LAYERS, sizeof(psnode_codec));'
           'return 0;}\n')
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, "s.c"), os.path.join(td, "s")
        open(cfile, "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(q) for q in out]
    assert sizes[:4] == [C.sizeof(_native.Mlp), C.sizeof(_native.Series), C.sizeof(_native.Problem), C.sizeof(_native.Adjoint)]
    assert sizes[4] == _native.IMPL_TC and sizes[5] == _native.PSNODE_MAX_LAYERS
    assert sizes[6] == C.sizeof(_native.Codec)


def test_param_count_helper(native_lib):
    from py_psnode_b200 import _native
    m = _native.Mlp()
    m.n_layers = 2
    m.in_dim[0], m.out_dim[0], m.in_dim[1], m.out_dim[1] = 54, 64, 64, 16
    assert native_lib.psnode_mlp_param_count(C.byref(m)) == 54 * 64 + 64 + 64 * 16 + 16


def test_product_never_imports_the_oracle():
    bad = []
    for base in ("py_psnode_b200", "neural_dae"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "psnode_oracle" in text:
                        bad.append(os.path.join(dirpath, f))
    for f in ("utils.py",):
        if "oracle" in open(os.path.join(ROOT, f)).read():
            bad.append(f)
    assert not bad, f"product code references the oracle: {bad}"


def test_cpu_tensors_fail_loudly():
    """integrate_ODE on CPU tensors must raise (no CPU fallback); so must a module the kernel cannot represent."""
    from py_psnode_b200 import DE_Func, RK4, UnsupportedModuleError
    de = DE_Func(x_dim=4, z_dim=1, hidden_dim=8)
    T, B = 5, 3
    t = torch.zeros(T, B, 1)
    x, z = torch.zeros(T, B, 4), torch.zeros(T, B, 1)
    a0 = torch.zeros(B, 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RK4().integrate_ODE(x_func=de, t=t, x=x, z=z, all_initial=a0)

    class Weird(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.x_dot = torch.nn.Sequential(torch.nn.Linear(15, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4))

        def forward(self, t0, xt, zt, all_initial):
            return xt
    with pytest.raises(UnsupportedModuleError):
        RK4().integrate_ODE(x_func=Weird(), t=t, x=x, z=z, all_initial=a0)
    with pytest.raises(ValueError):
        RK4(step_size=0.1, grid_constructor=lambda *a: None)


def test_fused_loss_has_no_cpu_path():
    """losses.masked_sse is CUDA-only like the integrators; parallel.masked_mse_sum keeps plain torch for host tensors
    (the gloo tests of the sharding logic)."""
    from py_psnode_b200 import parallel
    from py_psnode_b200.losses import masked_sse
    pred, target, mask = torch.randn(3, 4, 2), torch.randn(3, 4, 2), torch.ones(3, 4, 1)
    with pytest.raises(TypeError, match="no CPU path"):
        masked_sse(pred, target, mask)
    num, den = parallel.masked_mse_sum(pred, target, mask)
    assert torch.allclose(num, ((pred - target) ** 2 * mask).sum()) and float(den) == 12.0


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from py_psnode_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_native.NativeLibraryError):
        _native.lib()


def test_step_integrate_contract_on_cpu():
    """step_integrate is the reference's public single-step API with an arbitrary module (my_solvers.py:48-50)."""
    from py_psnode_b200 import Euler, Midpoint, RK4
    f = lambda t0, xt, zt, all_initial: -xt
    x0 = torch.ones(2, 3)
    dt = torch.full((2, 1), 0.1)
    for cls, want in ((Euler, 0.9), (Midpoint, 1 - 0.1 + 0.005), (RK4, 0.9048375)):
        x1, f0 = cls().step_integrate(func=f, t0=torch.zeros(2, 1), dt=dt, t1=dt, x0=x0, z0=None, all_initial=None)
        assert torch.allclose(x1, torch.full_like(x0, want), atol=1e-6)
        assert torch.equal(f0, -x0)
    assert (Euler.order, Midpoint.order, RK4.order) == (1, 2, 4)


def test_mma_descriptors_stay_in_uniform_registers(native_lib):
    """SASS check of the built objects: no R2UR (vector -> uniform register move) chain directly in front of the UTCHMMA
    of the product tensor-core kernels.  ptxas spilling the tile descriptors to vector registers cost 3.5 % at cfg3
    (DESIGN.md section 9); tools/sass_r2ur_check.py prints the same count per kernel."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    lib_dir = os.path.join(ROOT, "py_psnode_b200", "_lib")
    for unit, allowed in (("psnode_tc8_fwd.o", 0), ("psnode_tc_bwd.o", 2), ("psnode_tc_bwd_dae.o", 4), ("psnode_wide_fwd.o", 0),
                          ("psnode_wide_bwd.o", 0), ("psnode_wide_grad.o", 0), ("psnode_wide_proj.o", 0), ("psnode_lg.o", 1)):
        obj = os.path.join(lib_dir, unit)
        if not os.path.exists(obj):
            pytest.skip(f"{unit} not built")
        txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True, check=True).stdout
        kernels = re.split(r"\n\s*Function : ", txt)[1:]
        assert kernels, f"no kernels found in {unit}"
        for f in kernels:
            ops = [m.group(1) for m in (re.search(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", l) for l in f.split("\n")) if m]
            mma = [i for i, o in enumerate(ops) if "UTCHMMA" in o]
            if not mma:
                continue
            near = sum(1 for i, o in enumerate(ops) if "R2UR" in o and any(0 < j - i <= 14 for j in mma))
            assert near <= allowed, f"{unit}: {near} R2UR in front of UTCHMMA in {f.splitlines()[0][:80]}"


def test_blackwell_native_sass_of_the_latent_width_kernels(native_lib):
    """What proves the round-2 kernels are Blackwell-native (B200_PROFILING.md): tcgen05.mma -> UTCHMMA, tcgen05.ld / st -> LDTM / STTM,
    cp.async.bulk.tensor (the TMA-staged input series) -> UTMALDG, cp.async.bulk -> UBLKCP, in the objects that ship."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    lib_dir = os.path.join(ROOT, "py_psnode_b200", "_lib")
    want = {"psnode_wide_proj.o": ("UTCHMMA", "UTMALDG", "LDTM", "STTM"), "psnode_wide_fwd.o": ("UTCHMMA", "UBLKCP", "LDTM", "STTM"),
            "psnode_wide_bwd.o": ("UTCHMMA", "LDTM", "STTM"), "psnode_wide_grad.o": ("UTCHMMA", "UBLKCP", "LDTM"),
            "psnode_lg.o": ("UTCHMMA", "UTMALDG", "LDTM")}
    for unit, mnemonics in want.items():
        obj = os.path.join(lib_dir, unit)
        if not os.path.exists(obj):
            pytest.skip(f"{unit} not built")
        txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True, check=True).stdout
        for mn in mnemonics:
            assert mn in txt, f"{unit}: no {mn} in the SASS"


def test_round2_entries_have_no_cpu_path():
    """The fused model-pipeline entries (encoders / decoders, Init_Func) are CUDA-only like the rest of the path: CPU tensors raise."""
    import torch.nn as nn
    from py_psnode_b200 import AE_Func, DE_Func, RK4
    H, B, T = 128, 3, 4
    codec = lambda i, o: nn.Sequential(nn.Linear(i, H), nn.ELU(), nn.Linear(H, o))
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2)
    t = torch.zeros(T, B, 1)
    with pytest.raises((RuntimeError, TypeError), match="CUDA|cuda"):
        RK4().integrate_ODE_encoded(x_func=de, t=t, x0=torch.zeros(B, H), z=torch.zeros(T, B, 2), all_initial=torch.zeros(B, 2 * H),
                                    z_encoder=codec(2, H), x_decoder=codec(H, 3))
    init = nn.Module()
    init.init_fun = nn.Sequential(nn.Linear(5, 8), nn.ELU(), nn.Linear(8, 4))
    init.forward = lambda z0, v0, i0: init.init_fun(torch.cat([z0, v0, i0], dim=-1))
    with pytest.raises((RuntimeError, TypeError), match="CUDA|cuda"):
        RK4.init_state(init, torch.zeros(B, 1), torch.zeros(B, 2), torch.zeros(B, 2))


def test_codec_modules_are_pattern_checked():
    """Encoders / decoders must be nn.Sequential(Linear, ELU, Linear); anything else is rejected before any kernel runs."""
    import torch.nn as nn
    from py_psnode_b200 import pattern
    ok = nn.Sequential(nn.Linear(2, 8), nn.ELU(), nn.Linear(8, 8))
    assert len(pattern.codec_chain(ok, "z_encoder")) == 2
    for bad in (nn.Sequential(nn.Linear(2, 8), nn.ReLU(), nn.Linear(8, 8)), nn.Sequential(nn.Linear(2, 8)), nn.Linear(2, 8),
                nn.Sequential(nn.Linear(2, 8), nn.ELU(), nn.Linear(8, 8), nn.ELU(), nn.Linear(8, 8))):
        with pytest.raises(pattern.UnsupportedModuleError):
            pattern.codec_chain(bad, "z_encoder")
