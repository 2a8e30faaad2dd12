#!/bin/bash
# GPU call b: full -m gpu suite after the math/kernels changes, config C / D / E kernel timings, ncu of the config C and log-3D kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r2b_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_gputest.log
timeout 300 python tools/bench_configs.py --cases Cdev,Dk,Edev --reps 10 > gpurun_out/r2b_configs.jsonl 2> gpurun_out/r2b_configs.err
C="python tools/bench_configs.py --cases Cdev --reps 1"
timeout 200 $C > gpurun_out/r2b_plain_c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive2d_cta -s 3 -c 2 -f -o gpurun_out/r2b_configc $C > gpurun_out/r2b_ncu_c.log 2>&1
export FTS_BENCH_DK=log_3d:8
D="python tools/bench_configs.py --cases Dk --reps 1"
timeout 200 $D > gpurun_out/r2b_plain_d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_level_grid -s 2 -c 1 -f -o gpurun_out/r2b_log3d $D > gpurun_out/r2b_ncu_d.log 2>&1
tail -5 gpurun_out/r2b_gputest.log; cat gpurun_out/r2b_configs.jsonl | cut -c1-250
