"""bench.py's JSON-line contract, checked on the CPU without running a kernel: our arm's line is assembled from a run_workload() result by
`bench.assemble_line`, the reference arm prints the SAME workload-only `config` object (the driver compares them), every workload has the
per-unit figures the roofline uses, and the sharding rule gives each rank the batch BASELINE quotes."""
import argparse
import json

import bench


def _fake_result(name, w, world):
    B = bench.rank_batch(w, world)
    return {"workload": name + ": " + w["desc"], "batch_per_gpu": B, "grid_steps": w["N"], "scaling": w["scaling"], "value": 1.0e8,
            "unit": "traj-steps/s", "ms_per_step": 7.0, "kernel": "psn_tc8_ode_kernel<rk4>", "kernel_ms": 6.9, "gpu_launches": 10,
            "roofline": {"bound": "tensor", "achieved": 60.0, "peak": 1644.2, "unit": "TFLOP/s", "frac": 0.036, "traffic": None},
            "roofline_hbm": {"bound": "hbm", "achieved": 45.0, "peak": 6553.0, "unit": "GB/s", "frac": 0.0069, "traffic": None},
            "fp32": {}, "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []},
            "e2e": {"value": 9.0e7, "unit": "traj-steps/s", "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 2},
            "cpu_baseline": {"value": 2.0e6, "unit": "traj-steps/s", "cores": 16, "kind": "reference", "sample": "whole job"}}


def test_our_line_carries_every_contract_key_and_a_workload_only_config():
    for name, w in bench.WORKLOADS.items():
        for world in (1, 2, 4, 8):
            args = argparse.Namespace(steps=10, warmup=3, workload=name, gpus=world)
            line = bench.assemble_line(args, w, world, _fake_result(name, w, world), {"x": {"value": 1}}, "16 CPUs (0-15)")
            json.dumps(line)
            for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                      "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "kernel", "others"):
                assert k in line, (name, k)
            assert line["n_gpus"] == world and line["higher_is_better"] is True and line["vs_baseline"] is None
            assert line["config"] == bench.line_config(name, w, world)          # what `--impl reference` prints for the same run
            assert set(line["config"]) == {"workload", "batch_per_gpu", "global_batch", "grid_steps", "state_dim", "hidden", "parallelism", "l2"}
            assert line["config"]["workload"].startswith(name + ": ")
            for r in (line["roofline"], line["roofline_hbm"]):
                assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)


def test_workload_table_and_sharding():
    for name, w in bench.WORKLOADS.items():
        assert w["bytes_per_unit"] > 0 and w["flop_per_unit"] > 0 and w["scaling"] in ("weak", "strong"), name
        B, steps, note = bench.cpu_sample_plan(name, w)
        assert B == w["B"] and 1 <= steps <= w["N"] and note
    assert bench.rank_batch(bench.WORKLOADS["cfg2"], 8) == 4096                    # weak: per-GPU batch fixed
    assert bench.rank_batch(bench.WORKLOADS["cfg4"], 4) == 16384 // 4              # strong: BASELINE's 4-GPU shard
    assert bench.rank_batch(bench.WORKLOADS["cfg5"], 1) == 65536 // 8              # never fewer than 8 ways (what fits one GPU)
    assert bench.rank_batch(bench.WORKLOADS["cfg5"], 8) == 65536 // 8
    # reference-formulation FLOPs per traj-step of the 4-layer ODE_01 nets: 4 stages x 2 x (3S.H + 2 H.H + H.X)
    for name in ("cfg2", "ode01_h128"):
        w = bench.WORKLOADS[name]
        S, Hd, X = w["X"] + w["Z"], w["H"], w["X"]
        assert w["flop_per_unit"] == 4 * 2 * (3 * S * Hd + 2 * Hd * Hd + Hd * X), name
