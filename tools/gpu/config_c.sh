#!/bin/bash
# config C iteration: 2D tests, device-timed config C (loop + graph replay), ncu of the warp kernel.  usage: tools/gpu/config_c.sh TAG
TAG=${1:-c3}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "2d or 2D or config_c or sweep or kat or nvrtc or pinned" > gpurun_out/r2${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2${TAG}_gputest.log
tail -6 gpurun_out/r2${TAG}_gputest.log
timeout 300 python tools/bench_configs.py --cases Cdev --reps 10 > gpurun_out/r2${TAG}_configs.jsonl 2> gpurun_out/r2${TAG}_configs.err
cat gpurun_out/r2${TAG}_configs.jsonl; tail -3 gpurun_out/r2${TAG}_configs.err
C="python tools/bench_configs.py --cases Cdev --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive2d_warp -s 3 -c 1 -f -o gpurun_out/r2${TAG}_configc $C > gpurun_out/r2${TAG}_ncu_c.log 2>&1
tail -2 gpurun_out/r2${TAG}_ncu_c.log
