#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --permute --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
timeout 300 $B > gpurun_out/r2j_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_tail -s 3 -c 1 -f -o gpurun_out/r2j_bucket $B > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log
