#!/usr/bin/env python
"""bench.py -- RK4 trajectory-steps/s of the fused integrator on BASELINE.json's configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one `RK4().integrate_ODE(...)` call (the reference's hot path, neural_dae/my_solvers.py:52-80) over a
synthetic batch of B = 4096 trajectories x 1000 grid steps per GPU (ODE_01 `DE_Func`, X=16, Z=2,
H=64; t = 0.01*j; series ~ N(0, 0.1^2); seed 0; default nn.Linear init).  Batch sharding is the only parallelism
(independent trajectories, no data-path collective), so N GPUs integrate N x 4096 trajectories: "scaling": "weak".

Printed JSON (one line, rank 0): see the task contract.  `value` = traj-steps/s with inputs resident in HBM;
`e2e` = same metric through the public Python call with pinned-host inputs copied in and the trajectory copied back
inside the timed region; `roofline` = algorithmic HBM bytes / kernel time vs the measured copy bandwidth (the path is
FMA-bound, not HBM-bound: `fp32` gives the fraction of the fp32-FMA peak); `cpu_baseline` = the oracle port (same ATen
ops as the reference) timed on this box's host cores on a bounded sample.
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
    # name: (kind, X, Z, V, I, H, B per GPU, N steps, bytes/traj-step, FLOP/traj-step reference formulation)
    "cfg2": dict(kind="ode", X=16, Z=2, V=0, I=0, H=64, B=4096, N=1000, bytes_per_unit=76, flop_per_unit=101376,
                 desc="RK4 fixed-step, ODE_01 DE_Func 54-64-64-64-16 + external input z(t), batch 4096 x 1000 steps"),
    "cfg3": dict(kind="dae", X=16, Z=1, V=2, I=4, H=64, B=4096, N=1000, bytes_per_unit=96, flop_per_unit=131328,
                 desc="RK4 fixed-step, DAE_01 DE_Func 69-64-64-64-16 + AE_Func 42-64-64-64-4 (one explicit AE eval/step), "
                      "batch 4096 x 1000 steps"),
}
FP32_PEAK_TFLOPS = 74.4     # nominal: 148 SM x 128 lanes x 2 x 1.965 GHz; measured 72.1 by bench_micro/micro.cu


def make_problem(w, seed=0):
    """Synthetic inputs on the CPU (pinned by the caller when needed) + freshly initialised modules."""
    import torch
    from py_psnode_b200 import DE_Func, AE_Func
    torch.manual_seed(seed)
    B, T = w["B"], w["N"] + 1
    X, Z, V, I, H = w["X"], w["Z"], w["V"], w["I"], w["H"]
    de = DE_Func(x_dim=X, z_dim=Z, hidden_dim=H, v_dim=V, i_dim=I)
    ae = AE_Func(x_dim=X, v_dim=V, i_dim=I, hidden_dim=H, z_dim=Z) if w["kind"] == "dae" else None
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(T, 1, 1).repeat(1, B, 1).contiguous()
    g = torch.Generator().manual_seed(seed + 1)
    mk = lambda width: (torch.randn(T, B, width, generator=g) * 0.1)
    data = dict(t=t, z=mk(Z), x0=torch.randn(B, X, generator=g) * 0.1)
    if w["kind"] == "dae":
        data.update(v=mk(V), i0=torch.randn(B, I, generator=g) * 0.1)
    return de, ae, data


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


def cpu_baseline(w, sample_steps, repeats=3, threads=None):
    """The oracle port (same ATen ops as the reference's loop) on the host cores; returns traj-steps/s (best of `repeats`)."""
    import torch
    from oracle import psnode_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    de, ae, d = make_problem(w)
    T = sample_steps + 1
    t, z, x0 = d["t"][:T], d["z"][:T], d["x0"]
    B = t.shape[1]
    pd = [(m.weight.detach(), m.bias.detach()) for m in de.x_dot if hasattr(m, "weight")]
    x = x0.unsqueeze(0).expand(T, B, w["X"])
    best = float("inf")
    with torch.no_grad():
        for r in range(repeats + 1):          # first pass is the warm-up
            t0 = time.perf_counter()
            if w["kind"] == "ode":
                a0 = torch.cat((x0, z[0]), dim=-1)
                O.integrate_ode("rk4", pd, t, x, z, a0)
            else:
                pa = [(m.weight.detach(), m.bias.detach()) for m in ae.i_calculator if hasattr(m, "weight")]
                v, i0 = d["v"][:T], d["i0"]
                a0 = torch.cat((x0, z[0], v[0], i0), dim=-1)
                O.integrate_dae("rk4", pd, pa, x0, t, x, z, v, i0.unsqueeze(0).expand(T, B, w["I"]), a0)
            dt = time.perf_counter() - t0
            if r > 0:
                best = min(best, dt)
    return B * sample_steps / best, threads, best


def run_reference_arm(args, w, rank):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; /root/reference cannot travel to the
    GPU box and the reference has no installable package) on all host threads.  Each step = a bounded sample."""
    if rank != 0:
        return
    import torch
    sample_steps = 50
    threads = os.cpu_count()
    vals = []
    for k in range(args.warmup + args.steps):
        v, threads, secs = cpu_baseline(w, sample_steps, repeats=1, threads=threads)
        if k >= args.warmup:
            vals.append((v, secs))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(s for _, s in vals) / len(vals)
    sample = f"B={w['B']} x {sample_steps} RK4 steps per step (of {w['N']}), torch {torch.__version__} CPU, {threads} threads, no_grad"
    line = {"impl": "reference", "metric": "rk4_traj_steps_per_sec", "value": value, "unit": "traj-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": "traj-steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "traj-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "fused", "tc", "tc8"])
    ap.add_argument("--cpu-sample-steps", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the forward+reverse-sweep(+grad all-reduce) leg")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, w, rank)
        return

    import torch
    import torch.distributed as dist
    from py_psnode_b200 import RK4, _native
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the integration path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # each rank integrates its own shard of the global batch (seeded by rank): independent trajectories, no exchange
    de, ae, host = make_problem(w, seed=rank)
    de = de.to(dev)
    ae = ae.to(dev) if ae is not None else None
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    solver = RK4(impl=args.kernel)
    units = w["B"] * w["N"]

    # ---- kernel-resident timing ------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = call_integrate(w, solver, de, ae, resident)
        kernel_name = _native.last_kernel()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        sampler.mark_begin()
        launches0 = _native.launch_count()
        e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_all0.record()
        for a, b in evs:
            a.record()
            out = call_integrate(w, solver, de, ae, resident)
            b.record()
        e_all1.record()
        barrier()
        sampler.mark_end()
        launches = _native.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        total_ms = e_all0.elapsed_time(e_all1)
        per_call_ms = [a.elapsed_time(b) for a, b in evs]
    tt = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max = float(tt.item())
    ms_per_step = total_ms_max / args.steps
    value = units * world / (ms_per_step * 1e-3)
    kern_ms = statistics.mean(per_call_ms)      # dominant kernel ~ whole call (pack kernel is microseconds)

    # ---- end-to-end timing through the host-buffer C ABI (psnode_forward_host): inputs in pinned HOST memory, trajectory
    #      delivered to pinned HOST memory; every byte crosses PCIe inside the timed region (read / written in place by the
    #      kernel for pinned buffers, i.e. the copy is fused with the integration) -------------------------------------
    e2e = None
    if not args.no_e2e:
        import copy
        T = w["N"] + 1
        de_cpu = copy.deepcopy(de).cpu()
        ae_cpu = copy.deepcopy(ae).cpu() if ae is not None else None
        out_host = torch.empty((T, w["B"], w["X"]), dtype=torch.float32).pin_memory()
        iout_host = torch.empty((T, w["B"], w["I"]), dtype=torch.float32).pin_memory() if w["kind"] == "dae" else None
        moved = [0, 0]

        def e2e_step():
            d = pinned
            x_view = d["x0"].unsqueeze(0).expand(T, w["B"], w["X"])
            if w["kind"] == "ode":
                a0 = torch.cat((d["x0"], d["z"][0]), dim=-1)
                solver.integrate_ODE_host(x_func=de_cpu, t=d["t"], x=x_view, z=d["z"], all_initial=a0, out=out_host)
            else:
                i_view = d["i0"].unsqueeze(0).expand(T, w["B"], w["I"])
                a0 = torch.cat((d["x0"], d["z"][0], d["v"][0], d["i0"]), dim=-1)
                solver.integrate_DAE_host(x_init=d["x0"], x_func=de_cpu, i_func=ae_cpu, t=d["t"], x=x_view, z=d["z"], v=d["v"],
                                          i=i_view, all_initial=a0, out=(out_host, iout_host))
            moved[0], moved[1] = solver.last_host_bytes

        for _ in range(2):
            e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        a.record()
        for _ in range(args.steps):
            e2e_step()                                    # returns after the stream drained (results are in host memory)
        b.record()
        barrier()
        wall_ms = (time.perf_counter() - t_wall0) * 1e3
        ms = max(a.elapsed_time(b), wall_ms)              # the call blocks the host: take the larger of the two clocks
        tt = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item()) / args.steps
        e2e = {"value": units * world / (e2e_ms * 1e-3), "unit": "traj-steps/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": moved[0], "d2h_bytes_per_step": moved[1],
               "path": "psnode_forward_host (C ABI, HOST pointers): pinned buffers read/written in place over PCIe by the kernel"}

    # ---- training step: forward + reverse sweep (discrete adjoint) + ONE gradient all-reduce ---------------------
    train = None
    if not args.no_train:
        from py_psnode_b200 import parallel
        T = w["N"] + 1
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        x_target = torch.randn((T, w["B"], w["X"]), device=dev, generator=gen) * 0.1
        mask = torch.ones((T, w["B"], 1), device=dev)       # one value per (trajectory, grid point), as in the scripts
        i_target = torch.randn((T, w["B"], w["I"]), device=dev, generator=gen) * 0.1 if w["kind"] == "dae" else None
        plist = list(de.parameters()) + (list(ae.parameters()) if ae is not None else [])
        bucket = parallel.GradBucket(plist, n_extras=2)

        def numden(out):
            xs, is_ = out
            num, den = parallel.masked_mse_sum(xs, x_target, mask)
            if is_ is not None:
                num = num + parallel.masked_mse_sum(is_, i_target, mask)[0]
            return num, den

        def train_step():
            return parallel.sharded_training_step(lambda: call_integrate(w, solver, de, ae, resident), plist, bucket, numden)

        for _ in range(max(args.warmup, 3)):
            train_step()
        bwd_kernel = _native.last_kernel()
        barrier()
        l0 = _native.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            loss_val = train_step()
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tr_ms = float(tt.item()) / args.steps
        train = {"value": units * world / (tr_ms * 1e-3), "unit": "traj-steps/s", "ms_per_step": tr_ms,
                 "what": "forward + reverse sweep (discrete adjoint, all parameter grads) + masked-MSE + one flat gradient all-reduce",
                 "allreduce_bytes": bucket.nbytes, "kernel": bwd_kernel, "gpu_launches": int(_native.launch_count() - l0),
                 "loss": loss_val,
                 # forward + exact reverse mode = 3x the forward's algorithmic FLOPs (the tape-based sweep does not recompute)
                 "achieved_tflops_reference_formulation": 3 * w["flop_per_unit"] * units / (tr_ms * 1e-3) / 1e12}

    if rank == 0:
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
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get(kernel_name)
        except Exception:
            pass
        tflops = w["flop_per_unit"] * units / (kern_ms * 1e-3) / 1e12
        hbm = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
               "traffic": traffic, "peak_source": peak_src,
               "note": "algorithmic (compulsory) bytes per traj-step x units / kernel time; the path is ~1300 FLOP/byte, i.e. compute/latency bound"}
        if kernel_name.startswith("psn_tc"):
            tc_peak = peaks.get("bf16_tflops", 1590.0)
            tc_src = ("measured (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)" if "bf16_tflops" in peaks
                      else "fallback 1590 TFLOP/s dense bf16 (B200_PROFILING.md)")
            roofline = {"bound": "tensor", "achieved": tflops, "peak": tc_peak, "unit": "TFLOP/s", "frac": tflops / tc_peak,
                        "traffic": traffic, "peak_source": tc_src,
                        "note": f"achieved = algorithmic FLOPs of the reference formulation ({w['flop_per_unit']} per traj-step at "
                                f"{args.workload}) / kernel time. The kernel runs tcgen05 kind::tf32 (dense peak = half the bf16 figure) and "
                                "needs 3 MMAs per product (3xTF32) to hold the reference's fp32 accuracy, with N = 16 trajectories per MMA "
                                "(4096 trajectories / 148 SMs): it is bound by the serial layer chain (>= 16000 dependent layers per "
                                "trajectory), not by tensor throughput"}
        else:
            roofline = hbm
        line = {
            "metric": "rk4_traj_steps_per_sec", "value": value, "unit": "traj-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "batch_per_gpu": w["B"], "global_batch": w["B"] * world,
                       "grid_steps": w["N"], "state_dim": w["X"], "hidden": w["H"], "parallelism": f"batch-shard x{world}",
                       "kernel": kernel_name, "l2": "working set per call (inputs 49 MB + trajectory 262 MB) exceeds the 126 MB L2"},
            "roofline": roofline, "roofline_hbm": hbm,
            "fp32": {"achieved_tflops_reference_formulation": tflops, "peak_tflops_nominal": FP32_PEAK_TFLOPS,
                     "frac": tflops / FP32_PEAK_TFLOPS, "note": "CUDA-core fp32 FMA peak, for scale"},
            "kernel_ms": kern_ms, "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if train is not None:
            line["train"] = train
        if not args.no_cpu and world == 1:      # the CPU baseline is a 1-GPU-run item (other ranks would compete for the host cores)
            v, threads, secs = cpu_baseline(w, args.cpu_sample_steps)
            line["cpu_baseline"] = {"value": v, "unit": "traj-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"B={w['B']} x {args.cpu_sample_steps} of {w['N']} RK4 steps (per-step cost is constant), "
                                              f"best of 3 after 1 warm-up, {secs:.2f} s, torch CPU no_grad"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
