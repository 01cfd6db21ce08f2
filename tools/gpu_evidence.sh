#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, bench lines, ncu launch list and one full capture of the
# top kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; echo "bench cfg2 rc=$?"
python bench.py --steps 5 --warmup 3 --workload cfg3 --no-cpu > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; echo "bench cfg3 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>&1; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-train > gpurun_out/${TAG}_ncu_launch_run.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:psn_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_tc_fwd_cfg2 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-train > gpurun_out/${TAG}_ncu_full_run.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_smoke.log | tail -3; cat gpurun_out/${TAG}_bench_cfg2.json
