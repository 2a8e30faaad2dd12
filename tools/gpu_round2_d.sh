#!/bin/bash
# GPU call d: PDL for the queue passes (headline + config C), full tests, headline re-profile
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r2d_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_gputest.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
FTS_B200_PDL=0 timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link > gpurun_out/r2d_bench_nopdl.json 2> gpurun_out/r2d_bench_nopdl.err
timeout 300 python tools/bench_configs.py --cases Cdev --reps 20 > gpurun_out/r2d_configs.jsonl 2> gpurun_out/r2d_configs.err
FTS_B200_PDL=0 timeout 300 python tools/bench_configs.py --cases Cdev --reps 20 >> gpurun_out/r2d_configs.jsonl 2>> gpurun_out/r2d_configs.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
timeout 300 $B > gpurun_out/r2d_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_tpi -s 3 -c 1 -f -o gpurun_out/r2d_headline $B > gpurun_out/r2d_ncu.log 2>&1
tail -5 gpurun_out/r2d_gputest.log; cat gpurun_out/r2d_configs.jsonl | cut -c1-300
