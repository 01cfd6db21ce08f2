#!/usr/bin/env python
"""bench.py -- RK4 trajectory-steps/s of the fused integrator on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5|ode01_h128] [--no-others]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

The JSON line's top level is BASELINE.json's headline: configs[1] ("cfg2": RK4, ODE_01 DE_Func X=16 Z=2 H=64, B = 4096 x 1000
steps per GPU, weak scaling -- independent trajectories, no data-path collective).  One "step" = one `RK4().integrate_ODE(...)`
call (the reference's hot path, neural_dae/my_solvers.py:52-80) over the synthetic batch of SURVEY.md 8d (t = 0.01 j, series
~ N(0, 0.1^2), seeded, default nn.Linear init).  `value` = traj-steps/s with inputs resident in HBM; `e2e` = the same metric
through the host-buffer C ABI with pinned-host inputs and outputs (every byte crosses PCIe inside the timed region);
`train` = forward + reverse sweep + masked MSE + the ONE flat gradient all-reduce; `roofline` / `cpu_baseline` per the contract.

`others` carries the same measurements for the remaining BASELINE configs in the same run, so the driver's default
invocation records them:
  cfg3  RK4 DAE_01 (DE_Func + AE_Func, H = 64), B = 4096 x 1000 per GPU (weak)
  cfg4  RK4 ODE_02 latent net X = Z = H = 128, GLOBAL batch 16384 x 500 steps sharded over the N ranks (strong; BASELINE quotes
        it on 4 GPUs), training leg = adjoint with latent-input gradients + 0.67 MB gradient all-reduce
  cfg5  RK4 DAE_02 latent net H = 256, GLOBAL batch 65536 x 2000 steps sharded over 8 ranks (BASELINE quotes it on 8 GPUs); with
        fewer than 8 ranks each rank integrates one 1/8 shard (B = 8192).  Forward: per-layer tcgen05 GEMM launches (impl = layer)
        over all 2000 steps; the reverse sweep recomputes every step on the same GEMM kernel (psn_lg_backward) and the training leg
        also covers all 2000 steps (131 GiB of HBM: series, targets, trajectories and latent-input gradients are 16.8 GB each).  Only the
        e2e leg (GBs of pinned host memory per series) integrates a bounded number of grid steps, stated in `sample`.
CPU legs: the UNMODIFIED reference from oracle/_ref (vendored by oracle/make_ref.py, executed by oracle/ref_runner.py in a
subprocess; `kind: "reference"`), or the oracle port when oracle/_ref is absent (`kind: "port"`).
"""
import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # bytes_per_unit / flop_per_unit: SURVEY.md 8d (compulsory HBM bytes and reference-formulation FLOPs per trajectory-step)
    "cfg2": dict(kind="ode", net="01", X=16, Z=2, V=0, I=0, H=64, B=4096, N=1000, scaling="weak", bytes_per_unit=76, flop_per_unit=101376,
                 desc="RK4 fixed-step, ODE_01 DE_Func 54-64-64-64-16 + external input z(t), batch 4096 x 1000 steps"),
    "cfg3": dict(kind="dae", net="01", X=16, Z=1, V=2, I=4, H=64, B=4096, N=1000, scaling="weak", bytes_per_unit=96, flop_per_unit=131328,
                 desc="RK4 fixed-step, DAE_01 DE_Func 69-64-64-64-16 + AE_Func 42-64-64-64-4 (one explicit AE eval/step), "
                      "batch 4096 x 1000 steps"),
    "cfg4": dict(kind="ode", net="02", X=128, Z=128, V=0, I=0, H=128, B=16384, N=500, scaling="strong", quoted_gpus=4,
                 bytes_per_unit=1028, flop_per_unit=917504,
                 desc="RK4 fixed-step, ODE_02 latent DE_Func 768-128-128 (x_dim=z_dim=hidden=128), global batch 16384 x 500 steps, "
                      "adjoint training with latent-input gradients"),
    "cfg5": dict(kind="dae", net="02", X=256, Z=256, V=256, I=256, H=256, B=65536, N=2000, scaling="strong", quoted_gpus=8,
                 bytes_per_unit=4100, flop_per_unit=7864320, aux_steps=200, train_steps=2000,
                 desc="RK4 fixed-step, DAE_02 latent DE_Func 3072-256-256 + AE_Func 1792-256-256, global batch 65536 x 2000 steps"),
    # not a BASELINE config: cfg2 at the training script's argparse default --hidden 128 (neural_00_ODE_01_no_encode.py:245-247;
    # VERDICT r01 item 9): forward on psn_wide4_fwd_kernel, tape-based reverse sweep on psn_wide4_bwd_kernel + psn_wide_grad_kernel
    "ode01_h128": dict(kind="ode", net="01", X=16, Z=2, V=0, I=0, H=128, B=4096, N=1000, scaling="weak", bytes_per_unit=76,
                       flop_per_unit=333824,
                       desc="RK4 fixed-step, ODE_01 DE_Func 54-128-128-128-16 (the script's default --hidden 128; not a BASELINE config) "
                            "+ external input z(t), batch 4096 x 1000 steps"),
}
FP32_PEAK_TFLOPS = 74.4     # nominal: 148 SM x 128 lanes x 2 x 1.965 GHz; measured 72.1 by bench_micro/micro.cu


def line_config(name, w, world):
    """The `config` object of the JSON line: the workload only, nothing arm-specific -- both arms print the same one."""
    B = rank_batch(w, world)
    return {"workload": name + ": " + w["desc"], "batch_per_gpu": B,
            "global_batch": B * world if w["scaling"] == "weak" else w["B"],
            "grid_steps": w["N"], "state_dim": w["X"], "hidden": w["H"], "parallelism": f"batch-shard x{world}",
            "l2": "working set per call (cfg2: inputs 49 MB + trajectory 262 MB) exceeds the 126 MB L2"}


def rank_batch(w, world):
    """Trajectories one rank integrates: weak workloads keep B per GPU, strong ones shard the global batch (cfg5 never
    fewer than 8 ways: one 1/8 shard per rank is what fits and what BASELINE quotes)."""
    if w["scaling"] == "weak":
        return w["B"]
    ways = max(world, 8) if w.get("quoted_gpus") == 8 else world
    return w["B"] // ways


def make_modules(w, seed=0):
    import torch
    from py_psnode_b200 import DE_Func, AE_Func
    torch.manual_seed(seed)
    depth = 4 if w["net"] == "01" else 2
    X, Z, V, I, H = w["X"], w["Z"], w["V"], w["I"], w["H"]
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I, depth=depth)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z, depth=depth) if w["kind"] == "dae" else None
    return de, ae


def make_data(w, B, steps, seed=0, device="cpu"):
    """Synthetic inputs of SURVEY.md 8d.  On the CPU the draw order matches oracle/ref_runner.py (same tensors on both arms)."""
    import torch
    T = steps + 1
    X, Z, V, I = w["X"], w["Z"], w["V"], w["I"]
    g = torch.Generator(device=device).manual_seed(seed + 1)
    t = (torch.arange(T, dtype=torch.float32, device=device) * 0.01).view(T, 1, 1).repeat(1, B, 1).contiguous()
    mk = lambda width: (torch.randn(T, B, width, generator=g, device=device) * 0.1)
    data = dict(t=t, z=mk(Z), x0=torch.randn(B, X, generator=g, device=device) * 0.1)
    if w["kind"] == "dae":
        data.update(v=mk(V), i0=torch.randn(B, I, generator=g, device=device) * 0.1)
    return data


def call_integrate(w, solver, de, ae, d):
    """One call of the hot path through the public API.  `d` holds device tensors."""
    import torch
    T = d["t"].shape[0]
    B = d["t"].shape[1]
    x_view = d["x0"].unsqueeze(0).expand(T, B, w["X"])        # only x[0] is read without teacher forcing
    if w["kind"] == "ode":
        a0 = torch.cat((d["x0"], d["z"][0]), dim=-1)
        return solver.integrate_ODE(x_func=de, t=d["t"], x=x_view, z=d["z"], all_initial=a0), None
    i_view = d["i0"].unsqueeze(0).expand(T, B, w["I"])
    a0 = torch.cat((d["x0"], d["z"][0], d["v"][0], d["i0"]), dim=-1)
    return solver.integrate_DAE(x_init=d["x0"], x_func=de, i_func=ae, t=d["t"], x=x_view, z=d["z"], v=d["v"], i=i_view,
                                all_initial=a0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 25 ms.  Started BEFORE the warm-up (NVML initialisation inside a
    50 ms timed region perturbed it); only the samples stamped inside [mark_begin, mark_end] are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [q.strip() for q in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    stamp = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                    rows.append((stamp, float(f[1]), float(f[2]), f[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        window = "timed region"
        if not inside:          # region shorter than one sampling period: the samples closest to it (warm-up just before)
            inside, window = rows[-4:], "nearest samples (region shorter than the sampling period)"
        sm, mx, reasons = [r[1] for r in inside], [r[2] for r in inside], set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------- CPU legs
def ref_available():
    return os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "src", "neural_dae", "my_solvers.py"))


def cpu_reference(w, B, sample_steps, repeats, threads=None, timeout=900):
    """The UNMODIFIED reference (oracle/_ref) in a subprocess on the host cores.  Returns (traj-steps/s of the best repeat,
    threads, [seconds of every repeat])."""
    threads = threads or os.cpu_count()
    job = dict(kind=w["kind"], net=w["net"], X=w["X"], Z=w["Z"], V=w["V"], I=w["I"], H=w["H"], B=B, steps=sample_steps, seed=0,
               repeats=repeats, threads=threads, method="rk4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "time", json.dumps(job)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=os.path.join(ROOT, "oracle"))
    if out.returncode != 0:
        raise RuntimeError("oracle/ref_runner.py failed: " + out.stderr[-400:])
    res = json.loads(out.stdout.strip().splitlines()[-1])
    return B * sample_steps / res["seconds"], res["threads"], res["all_seconds"]


def cpu_port(w, B, sample_steps, repeats=2, threads=None):
    """The oracle port (same ATen ops as the reference's loop) on the host cores; same return convention as cpu_reference."""
    import torch
    from oracle import psnode_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    de, ae = make_modules(w)
    d = make_data(w, B, sample_steps)
    T = sample_steps + 1
    t, z, x0 = d["t"], d["z"], d["x0"]
    pd = [(m.weight.detach(), m.bias.detach()) for m in de.x_dot if hasattr(m, "weight")]
    x = x0.unsqueeze(0).expand(T, B, w["X"])
    secs = []
    with torch.no_grad():
        for r in range(repeats + 1):          # first pass is the warm-up
            t0 = time.perf_counter()
            if w["kind"] == "ode":
                a0 = torch.cat((x0, z[0]), dim=-1)
                O.integrate_ode("rk4", pd, t, x, z, a0)
            else:
                pa = [(m.weight.detach(), m.bias.detach()) for m in ae.i_calculator if hasattr(m, "weight")]
                v, i0 = d["v"], d["i0"]
                a0 = torch.cat((x0, z[0], v[0], i0), dim=-1)
                O.integrate_dae("rk4", pd, pa, x0, t, x, z, v, i0.unsqueeze(0).expand(T, B, w["I"]), a0)
            if r > 0:
                secs.append(time.perf_counter() - t0)
    return B * sample_steps / min(secs), threads, secs


def cpu_leg(w, B, sample_steps, repeats):
    """(value, cores, kind, seconds list)"""
    if ref_available():
        try:
            v, th, secs = cpu_reference(w, B, sample_steps, repeats)
            return v, th, "reference", secs
        except Exception as exc:           # fall back to the port, and say so
            sys.stderr.write(f"bench.py: reference runner failed ({exc}); timing the oracle port instead\n")
    v, th, secs = cpu_port(w, B, sample_steps, repeats)
    return v, th, "port", secs


def cpu_sample_plan(name, w):
    """(B, steps, note) of the CPU sample per workload: the full job where it costs seconds, >= the stated number of steps at
    full batch otherwise (the loop has no step-dependent state: per-step cost is constant)."""
    if name in ("cfg2", "cfg3", "ode01_h128"):
        return w["B"], w["N"], f"the whole job: B={w['B']} x {w['N']} RK4 steps"
    if name == "cfg4":
        return w["B"], 50, f"B={w['B']} (full batch) x 50 of {w['N']} RK4 steps, scaled linearly"
    return w["B"], 10, f"B={w['B']} (full batch) x 10 of {w['N']} RK4 steps, scaled linearly (one step = 257 GMAC on the CPU)"


def run_reference_arm(args, name, w, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on all host threads.  Rank 0 only."""
    if rank != 0:
        return
    import torch
    B, steps_full, note = cpu_sample_plan(name, w)
    # bound the whole K + W run to a few minutes: probe 20 steps, then size the per-step sample
    _, threads, kind, probe = cpu_leg(w, B, min(20, steps_full), 1)
    per_step = min(probe) / min(20, steps_full)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    sample_steps = max(min(steps_full, int(budget / per_step)), min(20, steps_full))
    v, threads, kind, secs = cpu_leg(w, B, sample_steps, args.steps + args.warmup - 1 if args.steps + args.warmup > 1 else 1)
    secs = secs[-args.steps:]
    value = B * sample_steps / statistics.mean(secs)
    ms = 1e3 * statistics.mean(secs)
    sample = (f"B={B} x {sample_steps} RK4 steps per step (of {w['N']}), torch {torch.__version__} CPU, {threads} threads, no_grad, "
              + ("unmodified reference from oracle/_ref with its event callbacks" if kind == "reference" else "oracle port"))
    line = {"impl": "reference", "metric": "rk4_traj_steps_per_sec", "value": value, "unit": "traj-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": line_config(name, w, max(args.gpus, 1)),
            "cpu_baseline": {"value": value, "unit": "traj-steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "traj-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------- GPU legs
def _e2e_leg(w, legs, host, resident, solver, de, ae, B, n_steps, units, active, steps, barrier, reduce_max):
    import torch
    e2e = None
    solver_device = next(de.parameters()).device
    if "e2e" in legs and host is not None:
        import copy
        T = n_steps + 1
        pinned = {k: v.pin_memory() for k, v in host.items()}
        de_cpu = copy.deepcopy(de).cpu()
        ae_cpu = copy.deepcopy(ae).cpu() if ae is not None else None
        out_host = torch.empty((T, B, w["X"]), dtype=torch.float32).pin_memory()
        iout_host = torch.empty((T, B, w["I"]), dtype=torch.float32).pin_memory() if w["kind"] == "dae" else None
        moved = [0, 0]

        def e2e_step():
            d = pinned
            x_view = d["x0"].unsqueeze(0).expand(T, B, w["X"])
            if w["kind"] == "ode":
                a0 = torch.cat((d["x0"], d["z"][0]), dim=-1)
                solver.integrate_ODE_host(x_func=de_cpu, t=d["t"], x=x_view, z=d["z"], all_initial=a0, out=out_host)
            else:
                i_view = d["i0"].unsqueeze(0).expand(T, B, w["I"])
                a0 = torch.cat((d["x0"], d["z"][0], d["v"][0], d["i0"]), dim=-1)
                solver.integrate_DAE_host(x_init=d["x0"], x_func=de_cpu, i_func=ae_cpu, t=d["t"], x=x_view, z=d["z"], v=d["v"],
                                          i=i_view, all_initial=a0, out=(out_host, iout_host))
            moved[0], moved[1] = solver.last_host_bytes

        # two transfer modes of the same C entry point, A/B'd in every run (the faster one is `e2e`):
        #   inplace : the kernel reads / writes the pinned buffers over PCIe itself (zero copy, fused with the integration)
        #   dma     : 8 time chunks, copy engines move chunk c+1's inputs and chunk c-1's trajectory rows while chunk c integrates
        timings = {}
        for mode in ("inplace", "dma"):
            if mode == "dma":
                os.environ["PSNODE_HOST_PATH"] = "dma"
            else:
                os.environ.pop("PSNODE_HOST_PATH", None)
            for _ in range(2):
                e2e_step()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_wall0 = time.perf_counter()
            a.record()
            for _ in range(steps):
                e2e_step()                                    # returns after the stream drained (results are in host memory)
            b.record()
            barrier()
            wall_ms = (time.perf_counter() - t_wall0) * 1e3
            timings[mode] = (reduce_max(max(a.elapsed_time(b), wall_ms)) / steps, moved[0], moved[1])   # the call blocks the host: larger clock
        os.environ.pop("PSNODE_HOST_PATH", None)
        # platform floor for this leg: the same bytes moved by plain pinned <-> device copies on two streams, all ranks at once
        # (no integration at all).  At 8 GPUs this is what bounds e2e: the ranks share the host's PCIe / memory bandwidth.
        copy_ms = None
        try:
            dev_in = {k: torch.empty_like(v, device=solver_device) for k, v in pinned.items()}
            dev_out = torch.empty_like(out_host, device=solver_device)
            dev_iout = torch.empty_like(iout_host, device=solver_device) if iout_host is not None else None
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

            def copy_step():
                with torch.cuda.stream(s_in):
                    for k in pinned:
                        dev_in[k].copy_(pinned[k], non_blocking=True)
                with torch.cuda.stream(s_out):
                    out_host.copy_(dev_out, non_blocking=True)
                    if dev_iout is not None:
                        iout_host.copy_(dev_iout, non_blocking=True)
                torch.cuda.synchronize()

            copy_step()
            barrier()
            t0c = time.perf_counter()
            for _ in range(steps):
                copy_step()
            barrier()
            copy_ms = reduce_max((time.perf_counter() - t0c) * 1e3) / steps
            del dev_in, dev_out, dev_iout
        except Exception:
            copy_ms = None
        best = min(timings, key=lambda k: timings[k][0])
        e2e_ms = timings[best][0]
        e2e = {"value": units * active / (e2e_ms * 1e-3), "unit": "traj-steps/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": timings[best][1], "d2h_bytes_per_step": timings[best][2],
               "path": "psnode_forward_host (C ABI, HOST pointers), mode " + best,
               "modes_ms": {k: v[0] for k, v in timings.items()},
               "copy_only_ms": copy_ms,
               "modes": "inplace = pinned buffers read/written over PCIe by the kernel itself; dma = chunked cudaMemcpyAsync on two "
                        "side streams overlapped with the integration (PSNODE_HOST_PATH=dma)"}
        del pinned, out_host, iout_host
    elif "e2e" in legs:
        # GB-sized series: staged end-to-end = pinned host -> device copy of the inputs, integrate, device -> pinned host copy
        T = n_steps + 1
        keys = [k for k in resident]
        pinned = {k: torch.empty(resident[k].shape, dtype=torch.float32).pin_memory() for k in keys}
        for k in keys:
            pinned[k].copy_(resident[k])
        out_host = torch.empty((T, B, w["X"]), dtype=torch.float32).pin_memory()
        iout_host = torch.empty((T, B, w["I"]), dtype=torch.float32).pin_memory() if w["kind"] == "dae" else None
        h2d = sum(v.numel() * 4 for v in pinned.values())
        d2h = out_host.numel() * 4 + (iout_host.numel() * 4 if iout_host is not None else 0)

        def e2e_step():
            with torch.no_grad():
                d = {k: resident[k].copy_(pinned[k], non_blocking=True) for k in keys}
                xs, is_ = call_integrate(w, solver, de, ae, d)
                out_host.copy_(xs, non_blocking=True)
                if is_ is not None:
                    iout_host.copy_(is_, non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        barrier()
        e2e_ms = reduce_max((time.perf_counter() - t_wall0) * 1e3) / steps
        e2e = {"value": units * active / (e2e_ms * 1e-3), "unit": "traj-steps/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "path": "pinned host -> cudaMemcpyAsync -> solver.integrate_* (device) -> cudaMemcpyAsync -> pinned host"}
        del pinned, out_host, iout_host
    return e2e


def _encoded_leg(w, solver, de, ae, resident, B, n_steps, units, active, barrier, reduce_max):
    """integrate_{ODE,DAE}_encoded at the workload's batch and ALL grid steps: raw (T,B,<=2) series in, decoded trajectories out; no
    (T,B,H) latent tensor exists (encoders inside the projection GEMMs, time chunks, decoders before the store)."""
    import torch
    import torch.nn as nn
    from py_psnode_b200 import _native
    dev = resident["t"].device
    H, T = w["H"], n_steps + 1
    dae = w["kind"] == "dae"
    XR, ZR, VR, IR = (32, 1, 2, 2) if dae else (8, 2, 0, 0)          # physical widths of SURVEY 8d (cfg5 / cfg4)
    torch.manual_seed(7)
    codec = lambda i, o: nn.Sequential(nn.Linear(i, H), nn.ELU(), nn.Linear(H, o)).to(dev)
    z_enc, x_dec = codec(ZR, H), codec(H, XR)
    v_enc, i_dec = (codec(VR, H), codec(H, IR)) if dae else (None, None)
    z_raw = torch.randn(T, B, ZR, device=dev) * 0.1
    v_raw = torch.randn(T, B, VR, device=dev) * 0.1 if dae else None
    x0 = resident["x0"]
    with torch.no_grad():
        a0 = torch.cat((x0, z_enc(z_raw[0])) + ((v_enc(v_raw[0]), resident["i0"]) if dae else ()), dim=-1)

    def call():
        if dae:
            return solver.integrate_DAE_encoded(x_init=x0, x_func=de, i_func=ae, t=resident["t"], z=z_raw, v=v_raw, all_initial=a0, z_encoder=z_enc,
                                                v_encoder=v_enc, x_decoder=x_dec, i_decoder=i_dec)
        return solver.integrate_ODE_encoded(x_func=de, t=resident["t"], x0=x0, z=z_raw, all_initial=a0, z_encoder=z_enc, x_decoder=x_dec)
    from py_psnode_b200 import engine
    torch.cuda.synchronize()
    engine._workspaces.clear()                       # so that the peak below includes this entry's own workspace, not a larger cached one
    torch.cuda.empty_cache()
    base = torch.cuda.memory_allocated()
    call()
    call()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        out = None
        out = call()
    b.record()
    barrier()
    ms = reduce_max(a.elapsed_time(b)) / 2
    del out
    peak_gib = (torch.cuda.max_memory_allocated() - base) / 2 ** 30
    # end to end with HOST buffers: the raw series come from pinned host memory every step and the decoded trajectories go back to it --
    # what a user of ODE_Model / DAE_Model.forward moves, ~100x fewer bytes than the latent series of the unfused calls
    zr_host = z_raw.cpu().pin_memory()
    vr_host = v_raw.cpu().pin_memory() if dae else None
    xo_host = torch.empty((T, B, XR), dtype=torch.float32).pin_memory()
    io_host = torch.empty((T, B, IR), dtype=torch.float32).pin_memory() if dae else None
    h2d = zr_host.numel() * 4 + (vr_host.numel() * 4 if dae else 0)
    d2h = xo_host.numel() * 4 + (io_host.numel() * 4 if dae else 0)

    def e2e_step():
        z_raw.copy_(zr_host, non_blocking=True)
        if dae:
            v_raw.copy_(vr_host, non_blocking=True)
        o = call()
        if dae:
            xo_host.copy_(o[0], non_blocking=True)
            io_host.copy_(o[1], non_blocking=True)
        else:
            xo_host.copy_(o, non_blocking=True)
        torch.cuda.synchronize()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        e2e_step()
    barrier()
    e2e_ms = reduce_max((time.perf_counter() - t0) * 1e3) / 2
    del zr_host, vr_host, xo_host, io_host
    return {"value": units * active / (ms * 1e-3), "unit": "traj-steps/s", "ms_per_step": ms, "kernel": _native.last_kernel(),
            "e2e": {"value": units * active / (e2e_ms * 1e-3), "unit": "traj-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "path": "pinned host raw series -> cudaMemcpyAsync -> integrate_*_encoded -> cudaMemcpyAsync -> pinned host "
                                                         "decoded trajectories, all grid steps"},
            "raw_widths": {"x": XR, "z": ZR, "v": VR, "i": IR},
            "peak_hbm_gib_above_inputs": peak_gib,
            "latent_series_gib_unfused": (4 if dae else 2) * T * B * H * 4 / 2 ** 30,
            "note": ("latent ODE_02 net: the wide kernels with the projection tiles generated from the raw series (psn_wide_forward_encoded) + the "
                     "decoder over the latent scratch; END TO END it moves the raw series and the decoded trajectories over PCIe instead of the "
                     "64x wider latent ones (compare `e2e` here with the workload's `e2e`)" if H == 128 else
                     "faster than torch encoders -> integrate_DAE -> torch decoders (740 ms at this shard), O(chunk) memory, and end to end it "
                     "covers ALL grid steps with 2.4 GB over PCIe (the unfused e2e leg needs 6.8 GB per 200 steps)"),
            "what": "Model.forward pipeline of the *_02 scripts in one call from the raw series: encoder hidden layers generated inside the hoisted "
                    "projection GEMMs, decoders before the store (psnode_forward_encoded); outputs decoded (T,B,x_dim)"}


def run_workload(name, w, args, ctx, steps, warmup, legs, main_line):
    """Measure one workload on this rank's GPU; returns the result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from py_psnode_b200 import RK4, _native
    rank, world, dev, local_rank = ctx["rank"], ctx["world"], ctx["dev"], ctx["local_rank"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        tt = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    B = rank_batch(w, world)
    n_steps = w["N"]                                      # grid steps per timed forward call: always the whole grid
    aux_steps = w.get("aux_steps", n_steps)               # e2e / training legs of cfg5: a bounded number of steps (stated)
    units = B * n_steps
    ways = w["B"] // B if w["scaling"] == "strong" else world
    active = min(world, ways)
    de, ae = make_modules(w, seed=0 if w["scaling"] == "strong" else rank)
    de = de.to(dev)
    ae = ae.to(dev) if ae is not None else None
    big = B * (n_steps + 1) * max(w["Z"], 1) * 4 > (1 << 30)
    host = None
    if big:                                               # GB-sized series are drawn on the device
        resident = make_data(w, B, n_steps, seed=rank, device=dev)
    else:
        host = make_data(w, B, n_steps, seed=rank)
        resident = {k: v.to(dev) for k, v in host.items()}
    solver = RK4(impl=args.kernel if main_line else "auto")
    res = {"workload": name + ": " + w["desc"], "batch_per_gpu": B, "grid_steps": n_steps, "scaling": w["scaling"]}
    if aux_steps != n_steps:
        res["sample"] = (f"`value` integrates all {n_steps} grid steps; the e2e leg integrates {aux_steps} steps per call (pinned host buffers of "
                         f"{aux_steps + 1} x {B} x {w['Z']} floats per series), the training leg {w.get('train_steps', aux_steps)} steps per call; "
                         "throughputs are per traj-step")
    if w["scaling"] == "strong":
        res["global_batch"] = w["B"]
        res["shards"] = ways
        if world < ways:
            res["note"] = f"{world} rank(s) each integrate one 1/{ways} shard of the global batch (BASELINE quotes this config on {w.get('quoted_gpus')} GPUs)"

    # ---- kernel-resident timing ------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0 and main_line:
        sampler.start()
    with torch.no_grad():
        out = None
        for _ in range(max(warmup, 3)):
            out = None                                     # free the previous trajectory first (cfg5: 2 x 16.8 GB per call)
            out = call_integrate(w, solver, de, ae, resident)
        kernel_name = _native.last_kernel()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        sampler.mark_begin()
        launches0 = _native.launch_count()
        e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_all0.record()
        for a, b in evs:
            out = None
            a.record()
            out = call_integrate(w, solver, de, ae, resident)
            b.record()
        e_all1.record()
        barrier()
        sampler.mark_end()
        launches = _native.launch_count() - launches0
        clocks = sampler.stop() if (rank == 0 and main_line) else None
        total_ms = e_all0.elapsed_time(e_all1)
        per_call_ms = [a.elapsed_time(b) for a, b in evs]
    del out
    ms_per_step = reduce_max(total_ms) / steps
    value = units * active / (ms_per_step * 1e-3)
    kern_ms = statistics.mean(per_call_ms)
    res.update(value=value, unit="traj-steps/s", ms_per_step=ms_per_step, kernel=kernel_name, kernel_ms=kern_ms,
               gpu_launches=int(launches))

    # ---- `*_02` models: the whole Model.forward pipeline from the RAW series (encoders + integration + decoders fused) ----
    if w["net"] == "02" and "e2e" in legs:
        try:
            res["encoded"] = _encoded_leg(w, solver, de, ae, resident, B, n_steps, units, active, barrier, reduce_max)
        except Exception as exc:
            res["encoded"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
            torch.cuda.synchronize()

    # ---- end-to-end through the host-buffer C ABI (psnode_forward_host): pinned HOST inputs and outputs ------------------
    e2e = None
    aux_res, aux_host, aux_units = resident, host, units
    if aux_steps != n_steps:
        del resident
        torch.cuda.empty_cache()
        aux_host = make_data(w, B, aux_steps, seed=rank)
        aux_res = {k: v.to(dev) for k, v in aux_host.items()}
        aux_units = B * aux_steps
    try:
        e2e = _e2e_leg(w, legs, aux_host, aux_res, solver, de, ae, B, aux_steps, aux_units, active, steps, barrier, reduce_max)
    except Exception as exc:
        e2e = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        torch.cuda.synchronize()

    # ---- training step: forward + reverse sweep (discrete adjoint) + ONE gradient all-reduce ---------------------
    train = None
    if "train" in legs:
        from py_psnode_b200 import parallel
        train_steps = w.get("train_steps", aux_steps)
        if train_steps != aux_steps:
            aux_res = aux_host = None
            torch.cuda.empty_cache()
            aux_res = make_data(w, B, train_steps, seed=rank, device=dev)
            aux_units = B * train_steps
        T = train_steps + 1
        memdbg = (lambda tag: print(f"[mem] {tag}: {torch.cuda.memory_allocated() / 2 ** 30:.1f} GiB allocated", file=sys.stderr, flush=True)) \
            if os.environ.get("PSNODE_BENCH_MEMDBG") else (lambda tag: None)
        memdbg("train leg, inputs ready")
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        x_target = torch.randn((T, B, w["X"]), device=dev, generator=gen) * 0.1
        mask = torch.ones((T, B, 1), device=dev)       # one value per (trajectory, grid point), as in the scripts
        i_target = torch.randn((T, B, w["I"]), device=dev, generator=gen) * 0.1 if w["kind"] == "dae" else None
        plist = list(de.parameters()) + (list(ae.parameters()) if ae is not None else [])
        bucket = parallel.GradBucket(plist, n_extras=2)
        tr_data = dict(aux_res)
        if w["net"] == "02":        # the latent input series are functions of the encoder weights: they carry gradients
            for k in ("z", "v"):
                if k in tr_data:
                    tr_data[k] = tr_data[k].detach().requires_grad_(True)

        def fused_forward():
            # integration fused with the scripts' masked-MSE numerator: the loss gradient is formed inside the reverse sweep,
            # dL/dx_sol (T,B,X) is never materialised
            d = tr_data
            Tt, Bt = d["t"].shape[0], d["t"].shape[1]
            x_view = d["x0"].unsqueeze(0).expand(Tt, Bt, w["X"])
            if w["kind"] == "ode":
                a0 = torch.cat((d["x0"], d["z"][0]), dim=-1)
                num, _ = solver.integrate_ODE_loss(x_func=de, t=d["t"], x=x_view, z=d["z"], all_initial=a0, target=x_target, mask=mask)
            else:
                i_view = d["i0"].unsqueeze(0).expand(Tt, Bt, w["I"])
                a0 = torch.cat((d["x0"], d["z"][0], d["v"][0], d["i0"]), dim=-1)
                num, _, _ = solver.integrate_DAE_loss(x_init=d["x0"], x_func=de, i_func=ae, t=d["t"], x=x_view, z=d["z"], v=d["v"], i=i_view,
                                                      all_initial=a0, target_x=x_target, target_i=i_target, mask=mask)
            memdbg("after the fused forward")
            return num, mask.sum()

        def train_step():
            for k in ("z", "v"):
                if k in tr_data and tr_data[k].requires_grad:
                    tr_data[k].grad = None
            memdbg("train step start")
            return parallel.sharded_training_step(fused_forward, plist, bucket, lambda out: out)

        torch.cuda.reset_peak_memory_stats()
        for _ in range(max(min(warmup, 3), 2)):
            train_step()
        bwd_kernel = _native.last_kernel()
        barrier()
        l0 = _native.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss_val = train_step()
        b.record()
        barrier()
        tr_ms = reduce_max(a.elapsed_time(b)) / steps
        train = {"value": aux_units * active / (tr_ms * 1e-3), "unit": "traj-steps/s", "ms_per_step": tr_ms,
                 "what": "forward + reverse sweep (discrete adjoint: all parameter grads" + (", latent-input grads" if w["net"] == "02" else "")
                         + ") with the masked-MSE loss fused into the sweep + one flat gradient all-reduce",
                 "allreduce_bytes": bucket.nbytes, "kernel": bwd_kernel, "gpu_launches": int(_native.launch_count() - l0),
                 "grid_steps": train_steps, "peak_hbm_gib": torch.cuda.max_memory_allocated() / 2 ** 30,
                 "loss": loss_val,
                 # forward + exact reverse mode = 3x the forward's algorithmic FLOPs (the tape-based sweeps do not recompute)
                 "achieved_tflops_reference_formulation": 3 * w["flop_per_unit"] * aux_units / (tr_ms * 1e-3) / 1e12}
        del x_target, mask, i_target, tr_data
        from py_psnode_b200 import engine
        engine.release_tape_pool()

    if rank != 0:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg_bytes = w["bytes_per_unit"] * units
    achieved_gbs = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {}).get(kernel_name)
    except Exception:
        pass
    tflops = w["flop_per_unit"] * units / (kern_ms * 1e-3) / 1e12
    hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
           "traffic": traffic, "peak_source": peak_src,
           "note": "algorithmic (compulsory) bytes per traj-step x units / call time; the path is ~900-1900 FLOP/byte, i.e. compute/latency bound"}
    if kernel_name.startswith("psn_tc") or kernel_name.startswith("psn_wide") or kernel_name.startswith("psn_lg"):
        tc_peak = peaks.get("bf16_tflops", 1590.0)
        tc_src = ("measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)" if "bf16_tflops" in peaks
                  else "fallback 1590 TFLOP/s dense bf16 (B200_PROFILING.md)")
        roofline = {"bound": "tensor", "achieved": tflops, "peak": tc_peak, "unit": "TFLOP/s", "frac": tflops / tc_peak,
                    "traffic": traffic, "peak_source": tc_src,
                    "note": f"achieved = algorithmic FLOPs of the reference formulation ({w['flop_per_unit']} per traj-step at {name}) / call "
                            "time. The kernels run tcgen05 kind::tf32 (dense peak = half the bf16 figure) with 3 MMAs per product (3xTF32) to "
                            "hold the reference's fp32 accuracy and N = 16 trajectories per MMA: bound by the serial layer chain, not by "
                            "tensor throughput"}
    else:
        roofline = hbm
    res.update(roofline=roofline, roofline_hbm=hbm,
               fp32={"achieved_tflops_reference_formulation": tflops, "peak_tflops_nominal": FP32_PEAK_TFLOPS,
                     "frac": tflops / FP32_PEAK_TFLOPS, "note": "CUDA-core fp32 FMA peak, for scale"},
               clocks=clocks)
    if e2e is not None:
        res["e2e"] = e2e
    if train is not None:
        res["train"] = train
    if "cpu" in legs and world == 1:      # the CPU baseline is a 1-GPU-run item (other ranks would compete for the host cores)
        cb, csteps, note = cpu_sample_plan(name, w)
        try:
            v, threads, kind, secs = cpu_leg(w, cb, csteps, 2 if name in ("cfg2", "cfg3") else 1)
            res["cpu_baseline"] = {"value": v, "unit": "traj-steps/s", "cores": threads, "kind": kind,
                                   "sample": f"{note}; best of {len(secs)} after 1 warm-up ({min(secs):.2f} s), torch CPU no_grad"}
        except Exception as exc:
            res["cpu_baseline"] = {"error": str(exc)[:300]}
    return res


def assemble_line(args, w, world, res, others, affinity):
    """The ONE JSON line of our arm from a run_workload() result (kept apart from main() so the contract's keys are testable on the CPU)."""
    line = {
        "metric": "rk4_traj_steps_per_sec", "value": res["value"], "unit": "traj-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": line_config(args.workload, w, world),
        "kernel": res["kernel"],
        "roofline": res["roofline"], "roofline_hbm": res["roofline_hbm"], "fp32": res["fp32"],
        "kernel_ms": res["kernel_ms"], "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
        "host_affinity": affinity,
    }
    for k in ("e2e", "train", "encoded", "cpu_baseline", "sample", "note"):
        if k in res:
            line[k] = res[k]
    if others:
        line["others"] = others
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "fused", "tc", "tc8", "wide"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the forward+reverse-sweep(+grad all-reduce) legs")
    ap.add_argument("--no-others", action="store_true", help="only the --workload line, no `others`")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, args.workload, w, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the integration path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # bind this rank to the CPUs next to its GPU before any pinned memory is touched: first-touch then places the pinned
    # batches on the GPU's NUMA node (the e2e legs move 0.3 - 8 GB per step over PCIe)
    affinity = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * k + b for k, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            affinity = f"{len(cpus)} CPUs ({cpus[0]}-{cpus[-1]})"
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    ctx = dict(rank=rank, world=world, dev=dev, local_rank=local_rank)
    legs = {"e2e", "train", "cpu"} - ({"e2e"} if args.no_e2e else set()) - ({"train"} if args.no_train else set()) \
        - ({"cpu"} if args.no_cpu else set())

    res = run_workload(args.workload, w, args, ctx, args.steps, args.warmup, legs, main_line=True)
    others = {}
    if not args.no_others:
        for name in WORKLOADS:
            if name == args.workload:
                continue
            try:
                torch.cuda.empty_cache()
                o = run_workload(name, WORKLOADS[name], args, ctx, steps=3, warmup=2, legs=legs, main_line=False)
            except Exception as exc:          # a failing extra workload must not take the headline line down
                o = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
                torch.cuda.synchronize()
            if rank == 0:
                others[name] = o

    if rank == 0:
        print(json.dumps(assemble_line(args, w, world, res, others, affinity)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
