"""Forward throughput and training-step time of Euler / Midpoint / RK4 at the cfg2 (ODE) and cfg3 (DAE) shapes.

    gpurun -- python tools/methods_probe.py
"""
import sys, time, torch
sys.path.insert(0, '.')
import bench
from py_psnode_b200 import Euler, Midpoint, RK4, _native
dev = torch.device("cuda:0")
for wl in ("cfg2", "cfg3"):
    w = bench.WORKLOADS[wl]
    de, ae, host = bench.make_problem(w)
    de = de.to(dev); ae = ae.to(dev) if ae is not None else None
    res = {k: v.to(dev) for k, v in host.items()}
    for S in (Euler, Midpoint, RK4):
        solver = S()
        with torch.no_grad():
            for _ in range(3): bench.call_integrate(w, solver, de, ae, res)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5): bench.call_integrate(w, solver, de, ae, res)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        # training step
        plist = list(de.parameters()) + (list(ae.parameters()) if ae is not None else [])
        def step():
            for p in plist: p.grad = None
            out = bench.call_integrate(w, solver, de, ae, res)
            loss = out[0].square().mean() + (out[1].square().mean() if out[1] is not None else 0)
            loss.backward()
        for _ in range(3): step()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): step()
        torch.cuda.synchronize(); dtr = (time.perf_counter() - t0) / 3
        print(f"{wl} {S.__name__}: fwd {dt*1e3:.2f} ms ({w['B']*w['N']/dt/1e6:.0f} M traj-steps/s), train step {dtr*1e3:.2f} ms, {_native.last_kernel()}", flush=True)
