#!/usr/bin/env python
"""Vendor the UNMODIFIED reference (xxh0523/Py_PSNODE @ d366e75) into the git-ignored `oracle/_ref/` so that it can travel to
the GPU box with the snapshot (git-ignored artefacts ship; `/root/reference` does not exist there).

    python oracle/make_ref.py            # run inside the build container (needs /root/reference)

    oracle/_ref/src/     byte-for-byte copies of the reference's Python files, original layout (NOT in git history)
    oracle/_ref/stubs/   stub packages for the two imports the reference needs and this image lacks:
                         `ray` (neural_dae/neural_base.py:4, an unused `from ray.worker import init`) and `matplotlib`
                         (plotting in the scripts' eval loops)

The reference is pure Python (no setup.py, nothing to compile), so "building" it is this copy.  It is TEST INFRASTRUCTURE and
the `--impl reference` arm of bench.py only: `oracle/ref_runner.py` executes it in a subprocess whose sys.path starts with
these two directories; the product path never imports it.
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

STUBS = {
    "ray/__init__.py": "from . import worker\n",
    "ray/worker.py": "def init(*a, **k):\n    return None\n",
    "matplotlib/__init__.py": "rcParams = {}\n\n\ndef use(*a, **k):\n    return None\n\n\nfrom . import pyplot, markers  # noqa: E402,F401\n",
    "matplotlib/pyplot.py": "def __getattr__(name):\n    raise AttributeError('matplotlib stub: plotting is not available (' + name + ')')\n",
    "matplotlib/markers.py": "",
}


def make(verbose: bool = True) -> str:
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"{REF} not found: oracle/_ref can only be (re)built inside the build container")
    src = os.path.join(DST, "src")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(src)
    n = 0
    for root, dirs, files in os.walk(REF):
        dirs[:] = [d for d in dirs if not d.startswith(".") and d != "__pycache__"]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, f), REF)
            os.makedirs(os.path.dirname(os.path.join(src, rel)), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), os.path.join(src, rel))
            n += 1
    for rel, text in STUBS.items():
        p = os.path.join(DST, "stubs", rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as fh:
            fh.write(text)
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("Unmodified copy of /root/reference (*.py) made by oracle/make_ref.py -- git-ignored, test infrastructure only.\n")
    if verbose:
        print(f"oracle/_ref: {n} reference files copied, {len(STUBS)} stub files written")
    return DST


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "neural_dae", "my_solvers.py"))


if __name__ == "__main__":
    make()
    sys.exit(0)
