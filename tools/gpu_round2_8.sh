#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
bash tools/gpu_round2_multi.sh 8 f > gpurun_out/r2f_8gpu_stdout.log 2>&1
bash tools/gpu_round2_multi.sh 4 g > gpurun_out/r2g_4gpu_stdout.log 2>&1
tail -c 800 gpurun_out/r2f_bench_8gpu.json; grep "^{" gpurun_out/r2f_hostlink_8gpu.jsonl | cut -c1-300
