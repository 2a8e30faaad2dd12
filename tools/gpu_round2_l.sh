#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --permute --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
timeout 300 $B > gpurun_out/r2l_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_adaptive1d|k_bucket" -c 16 --csv --log-file gpurun_out/r2l_launches.csv $B > gpurun_out/r2l_ncu.log 2>&1
grep -v "^==" gpurun_out/r2l_launches.csv | cut -d, -f5,15- | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_bucket -s 3 -c 1 -f -o gpurun_out/r2l_bucket $B > gpurun_out/r2l_ncu2.log 2>&1
tail -2 gpurun_out/r2l_ncu2.log
