#!/bin/bash
# first GPU call of round 2: full -m gpu suite (incl. the BASELINE-size parity tests), bench (both arms), ncu of the headline kernel
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
nproc > gpurun_out/r2a_nproc.txt; lscpu | head -20 >> gpurun_out/r2a_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gputest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
timeout 300 $B > gpurun_out/r2a_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_tpi -s 3 -c 1 -f -o gpurun_out/r2a_headline $B > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_gputest.log; tail -c 600 gpurun_out/r2a_bench.json
