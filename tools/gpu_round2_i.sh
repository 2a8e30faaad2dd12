#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --permute --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
timeout 300 $B > gpurun_out/r2i_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none -k regex:k_adaptive1d -c 12 --csv --log-file gpurun_out/r2i_launches.csv $B > gpurun_out/r2i_ncu.log 2>&1
grep -v "^==" gpurun_out/r2i_launches.csv | cut -d, -f5,12- | head -60
