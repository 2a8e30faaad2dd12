#!/bin/bash
# GPU call c: tests after the warp-shaped 2D kernel / re-entrant host entry points / integrate1D_cmpl, config C timings + ncu, headline re-profile
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r2c_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_gputest.log
timeout 300 python tools/bench_configs.py --cases Cdev,C --reps 20 > gpurun_out/r2c_configs.jsonl 2> gpurun_out/r2c_configs.err
FTS_B200_WARP2D_LEVELS=5 timeout 300 python tools/bench_configs.py --cases Cdev --reps 20 >> gpurun_out/r2c_configs.jsonl 2>> gpurun_out/r2c_configs.err
C="python tools/bench_configs.py --cases Cdev --reps 1"
timeout 200 $C > gpurun_out/r2c_plain_c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive2d -s 6 -c 2 -f -o gpurun_out/r2c_configc $C > gpurun_out/r2c_ncu_c.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
timeout 300 $B > gpurun_out/r2c_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_tpi -s 3 -c 1 -f -o gpurun_out/r2c_headline $B > gpurun_out/r2c_ncu.log 2>&1
tail -5 gpurun_out/r2c_gputest.log; cat gpurun_out/r2c_configs.jsonl | cut -c1-300
