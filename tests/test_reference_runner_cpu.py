"""oracle/ref_runner.py (the UNMODIFIED reference vendored in oracle/_ref, bench.py's `--impl reference` arm) integrates the same
synthetic problem as bench.py builds for the GPU arm and as the oracle port restates: same seeds -> same weights and inputs
-> the trajectory checksums agree to fp32 summation noise.  Skipped when oracle/_ref has not been vendored."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_available():
    return os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "src", "neural_dae", "my_solvers.py"))


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref not vendored (python oracle/make_ref.py, build container only)")
@pytest.mark.parametrize("name,B,steps", [("cfg2", 24, 12), ("cfg3", 16, 8), ("cfg4", 8, 5)])
def test_reference_runner_matches_oracle_port(name, B, steps):
    import bench
    from oracle import psnode_oracle as O
    w = bench.WORKLOADS[name]
    job = dict(kind=w["kind"], net=w["net"], X=w["X"], Z=w["Z"], V=w["V"], I=w["I"], H=w["H"], B=B, steps=steps, seed=0,
               repeats=0, threads=1, method="rk4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "time", json.dumps(job)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=os.path.join(ROOT, "oracle"))
    assert out.returncode == 0, out.stderr[-500:]
    ref = json.loads(out.stdout.strip().splitlines()[-1])
    torch.set_num_threads(1)
    de, ae = bench.make_modules(w)
    d = bench.make_data(w, B, steps)
    T = steps + 1
    pd = [(m.weight.detach(), m.bias.detach()) for m in de.x_dot if hasattr(m, "weight")]
    x = d["x0"].unsqueeze(0).expand(T, B, w["X"])
    with torch.no_grad():
        if w["kind"] == "ode":
            sol = O.integrate_ode("rk4", pd, d["t"], x, d["z"], torch.cat((d["x0"], d["z"][0]), dim=-1))
        else:
            pa = [(m.weight.detach(), m.bias.detach()) for m in ae.i_calculator if hasattr(m, "weight")]
            a0 = torch.cat((d["x0"], d["z"][0], d["v"][0], d["i0"]), dim=-1)
            sol = O.integrate_dae("rk4", pd, pa, d["x0"], d["t"], x, d["z"], d["v"], d["i0"].unsqueeze(0).expand(T, B, w["I"]), a0)[0]
    mine = float(sol.double().sum())
    assert abs(mine - ref["checksum"]) <= 1e-6 * max(1.0, abs(mine)), (mine, ref["checksum"])
