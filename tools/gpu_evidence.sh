#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, bench lines, ncu launch lists and full captures of the top
# kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; echo "bench cfg2 rc=$?"
python bench.py --steps 5 --warmup 3 --workload cfg3 --no-cpu --no-others > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; echo "bench cfg3 rc=$?"
# launch lists (forward-only bench command; training step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-train --no-others > gpurun_out/${TAG}_ncu_launch_run.log 2>&1; echo "ncu launches rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_train_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-others > gpurun_out/${TAG}_ncu_launch_train_run.log 2>&1; echo "ncu train launches rc=$?"
# full captures: forward (cfg2), reverse sweep (cfg2), DAE forward (cfg3)
ncu --set full --clock-control none --import-source on -k regex:psn_tc8 -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_tc8_fwd_cfg2 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-train --no-others > gpurun_out/${TAG}_ncu_full_fwd.log 2>&1; echo "ncu fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:psn_tc_bwd -s 1 -c 1 -f -o gpurun_out/${TAG}_ncu_tc_bwd_cfg2 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-others > gpurun_out/${TAG}_ncu_full_bwd.log 2>&1; echo "ncu bwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:psn_tc8 -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_tc8_dae_cfg3 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-train --no-others --workload cfg3 > gpurun_out/${TAG}_ncu_full_dae.log 2>&1; echo "ncu dae rc=$?"
ncu --set full --clock-control none --import-source on -k regex:psn_tc_bwd_dae -s 1 -c 1 -f -o gpurun_out/${TAG}_ncu_tc_bwd_dae_cfg3 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-others --workload cfg3 > gpurun_out/${TAG}_ncu_full_bwd_dae.log 2>&1; echo "ncu dae bwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:psn_masked_sse -s 6 -c 3 -f -o gpurun_out/${TAG}_ncu_masked_sse_cfg2 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-others > gpurun_out/${TAG}_ncu_full_loss.log 2>&1; echo "ncu loss rc=$?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench_cfg2.json
