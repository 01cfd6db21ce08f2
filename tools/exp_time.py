"""Time the cfg2 / cfg3 forward (inference and tape-writing) for each library variant:
    gpurun -- python tools/exp_time.py [cfg2|cfg3] [lib.so ...]      (no lib = the regular build)"""
import os, subprocess, sys
if os.environ.get("_EXP_CHILD"):
    sys.path.insert(0, ".")
    import torch, bench
    from py_psnode_b200 import RK4, _native, engine
    wl = os.environ["_EXP_WL"]
    w = bench.WORKLOADS[wl]
    dev = torch.device("cuda:0")
    de, ae, host = bench.make_problem(w)
    de = de.to(dev); ae = ae.to(dev) if ae is not None else None
    res = {k: v.to(dev) for k, v in host.items()}
    solver = RK4()
    def timed(fn, n=5):
        for _ in range(3): fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize(); ev[0].record()
        for _ in range(n): fn()
        ev[1].record(); torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / n
    with torch.no_grad():
        t_inf = timed(lambda: bench.call_integrate(w, solver, de, ae, res))
    k = _native.last_kernel()
    plist = list(de.parameters()) + (list(ae.parameters()) if ae is not None else [])
    def fwd_tape():
        out = bench.call_integrate(w, solver, de, ae, res)
        return out
    t_tape = timed(fwd_tape)
    k2 = _native.last_kernel()
    print(f"{os.environ.get('PSNODE_B200_LIB', 'default')}: {wl} inference {t_inf:.3f} ms [{k}]  with tape {t_tape:.3f} ms [{k2}]", flush=True)
else:
    args = sys.argv[1:]
    wl = args.pop(0) if args and args[0] in ("cfg2", "cfg3") else "cfg2"
    for lib in (args or [None]):
        env = dict(os.environ, _EXP_CHILD="1", _EXP_WL=wl)
        if lib: env["PSNODE_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, __file__], env=env)
