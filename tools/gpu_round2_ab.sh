#!/bin/bash
# GPU call ab: n-D tensor product throughput + ncu of k_fixed_nd
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py --cases ND --reps 10 > gpurun_out/r2ab_configs.jsonl 2> gpurun_out/r2ab_configs.err
cut -c1-400 gpurun_out/r2ab_configs.jsonl; tail -3 gpurun_out/r2ab_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fixed_nd -s 2 -c 1 -f -o gpurun_out/r2ab_nd python tools/bench_configs.py --cases ND --reps 2 > gpurun_out/r2ab_ncu.log 2>&1
tail -3 gpurun_out/r2ab_ncu.log
