#!/bin/bash
# ncu --set full capture of k_level_grid<double,F3Exp,3> at level 8 (config D's kernel).  usage: tools/gpu/ncu_exp3d.sh
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
D="python tools/bench_configs.py --cases Dk --reps 1"
FTS_BENCH_DK=exp_3d:8 timeout 200 $D > gpurun_out/r2az_plain_d.log 2>&1 && FTS_BENCH_DK=exp_3d:8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_level_grid -s 2 -c 1 -f -o gpurun_out/r2az_exp3d $D > gpurun_out/r2az_ncu_d.log 2>&1
tail -2 gpurun_out/r2az_ncu_d.log; cat gpurun_out/r2az_plain_d.log | cut -c1-200
