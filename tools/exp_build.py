"""Build A/B variants of one translation unit: python tools/exp_build.py psnode_tc8_fwd.cu 1 2 3  ->  _lib/exp/libv{N}.so
(the unit compiled with -DPSN_EXP=N, linked against the objects of the regular build)."""
import glob, os, subprocess, sys
sys.path.insert(0, ".")
from py_psnode_b200 import build as B
B.build()
unit, variants = sys.argv[1], sys.argv[2:]
out = os.path.join(B.LIB_DIR, "exp"); os.makedirs(out, exist_ok=True)
objs = [o for o in glob.glob(os.path.join(B.LIB_DIR, "*.o")) if os.path.basename(o) != unit[:-3] + ".o"]
procs = []
for v in variants:
    obj = os.path.join(out, f"v{v}.o")
    cmd = [B._nvcc(), *[f for f in B.NVCC_FLAGS if f != "-shared"], f"-DPSN_EXP={v}", "-c", os.path.join(B.CSRC, unit), "-o", obj]
    procs.append((v, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for v, obj, p in procs:
    o, _ = p.communicate()
    if p.returncode: raise SystemExit(o)
    lib = os.path.join(out, f"libv{v}.so")
    r = subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", lib, obj, *objs],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode: raise SystemExit(r.stdout)
    print("built", lib)
