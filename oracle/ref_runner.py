#!/usr/bin/env python
"""Run the UNMODIFIED reference from oracle/_ref (see oracle/make_ref.py) in THIS process.  Always started as a subprocess
(`python oracle/ref_runner.py ...`) by bench.py's CPU legs and by tests, because the reference's package is also called
`neural_dae` and must not be imported next to the repo's shim.  Test infrastructure: nothing under py_psnode_b200/ uses it.

    python oracle/ref_runner.py time '<json workload>'      -> one JSON line {"seconds": best, "units": B*steps, ...}

Workload JSON: {"kind": "ode"|"dae", "net": "01"|"02", "X","Z","V","I","H", "B", "steps", "seed", "repeats", "threads",
"method": "rk4"|"euler"|"midpoint"}.  Inputs follow SURVEY.md 8d (t = 0.01 j, series ~ N(0, 0.1^2), seeded) and are
generated exactly as bench.py's make_problem does; the solver object is the reference's own (neural_dae/my_fixed_grid.py),
the RHS modules are the script-local DE_Func / AE_Func classes, the event callbacks are passed as the scripts pass them
(event_t = -5: the per-step predicate runs and never fires).
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, "_ref", "src")
REF_STUBS = os.path.join(HERE, "_ref", "stubs")


def load_reference():
    if not os.path.isfile(os.path.join(REF_SRC, "neural_dae", "my_solvers.py")):
        raise SystemExit("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container")
    repo_root = os.path.dirname(HERE)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (repo_root, HERE)]
    sys.path.insert(0, REF_SRC)
    sys.path.insert(0, REF_STUBS)
    import importlib
    nd = importlib.import_module("neural_dae")
    assert os.path.abspath(nd.__file__).startswith(REF_SRC), nd.__file__
    return nd, importlib


def synth(w, torch):
    """Same tensors as bench.py make_problem (generator seeded with seed + 1, same draw order)."""
    B, T = w["B"], w["steps"] + 1
    g = torch.Generator().manual_seed(w.get("seed", 0) + 1)
    full_T = w.get("full_T", T)                     # bench draws the whole series and slices: keep the same stream
    t = (torch.arange(full_T, dtype=torch.float32) * 0.01).view(full_T, 1, 1).repeat(1, B, 1)[:T].contiguous()
    mk = lambda width: (torch.randn(full_T, B, width, generator=g) * 0.1)[:T]
    d = dict(t=t, z=mk(w["Z"]), x0=torch.randn(B, w["X"], generator=g) * 0.1)
    if w["kind"] == "dae":
        d.update(v=mk(w["V"]), i0=torch.randn(B, w["I"], generator=g) * 0.1)
    return d


def cmd_time(w):
    import torch
    nd, importlib = load_reference()
    threads = int(w.get("threads") or os.cpu_count())
    torch.set_num_threads(threads)
    torch.manual_seed(w.get("seed", 0))
    name = {("ode", "01"): "neural_00_ODE_01_no_encode", ("ode", "02"): "neural_00_ODE_02_direct_encode",
            ("dae", "01"): "neural_01_DAE_01_no_encode", ("dae", "02"): "neural_01_DAE_02_direct_encode"}[(w["kind"], w["net"])]
    mod = importlib.import_module(name)
    X, Z, V, I, H = w["X"], w["Z"], w["V"], w["I"], w["H"]
    if w["kind"] == "ode":
        de = mod.DE_Func(x_dim=X, z_dim=Z, hidden_dim=H)
        ae = None
        ev = nd.ODE_Event()
    else:
        de = mod.DE_Func(x_dim=X, z_dim=Z, v_dim=V, i_dim=I, hidden_dim=H)
        ae = mod.AE_Func(x_dim=X, z_dim=Z, v_dim=V, i_dim=I, hidden_dim=H)
        ev = nd.DAE_Event()
    d = synth(w, torch)
    T, B = d["t"].shape[0], d["t"].shape[1]
    solver = {"rk4": nd.RK4, "euler": nd.Euler, "midpoint": nd.Midpoint}[w.get("method", "rk4")]()
    event_t = torch.full((B, 1, 1), -5.0)
    x = d["x0"].unsqueeze(0).expand(T, B, X)
    best = float("inf")
    every = []
    with torch.no_grad():
        for r in range(int(w.get("repeats", 1)) + 1):        # first pass = warm-up
            t0 = time.perf_counter()
            if w["kind"] == "ode":
                ev.set_event(t=event_t, z=torch.zeros(B, 1, Z))
                a0 = torch.cat((d["x0"], d["z"][0]), dim=-1)
                out = solver.integrate_ODE(x_func=de, t=d["t"], x=x, z=d["z"], all_initial=a0, event_fn=ev.event_fn,
                                           jump_change_fn=ev.jump_change_fn)
            else:
                ev.set_event(t=event_t, z=torch.zeros(B, 1, Z), v=torch.zeros(B, 1, V))
                a0 = torch.cat((d["x0"], d["z"][0], d["v"][0], d["i0"]), dim=-1)
                out = solver.integrate_DAE(x_init=d["x0"], x_func=de, i_func=ae, t=d["t"], x=x, z=d["z"], v=d["v"],
                                           i=d["i0"].unsqueeze(0).expand(T, B, I), all_initial=a0, event_fn=ev.event_fn,
                                           jump_change_fn=ev.jump_change_fn)[0]
            dt = time.perf_counter() - t0
            if r > 0 or int(w.get("repeats", 1)) == 0:
                best = min(best, dt)
                every.append(dt)
    print(json.dumps({"seconds": best, "all_seconds": every, "units": B * w["steps"], "threads": threads, "torch": torch.__version__,
                      "checksum": float(out.double().sum())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] != "time":
        raise SystemExit(__doc__)
    cmd_time(json.loads(sys.argv[2]))
