#!/bin/bash
# config C iteration (short): device-timed config C incl. graph replay.  usage: gpu_round2_c2.sh TAG
TAG=${1:-c2}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py --cases Cdev --reps 10 > gpurun_out/r2${TAG}_configs.jsonl 2> gpurun_out/r2${TAG}_configs.err
cat gpurun_out/r2${TAG}_configs.jsonl; tail -5 gpurun_out/r2${TAG}_configs.err
FTS_BENCH_C_MAX_LEVELS=4 timeout 300 python tools/bench_configs.py --cases Cdev --reps 10 > gpurun_out/r2${TAG}_configs_ml4.jsonl 2> gpurun_out/r2${TAG}_configs_ml4.err
cat gpurun_out/r2${TAG}_configs_ml4.jsonl
