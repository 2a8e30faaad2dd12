#!/bin/bash
# config C kernel at 1x .. 32x BASELINE's batch (4096 boxes = one wave of warps): device time per call from the CUDA-graph replay.
# usage: tools/gpu/config_c_batch_sweep.sh TAG
TAG=${1:-cs}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
: > gpurun_out/r2${TAG}_config_c_batch_sweep.jsonl
for n in 1024 2048 4096 8192 16384 32768 65536 131072; do
  FTS_BENCH_C_BATCH=$n timeout 200 python tools/bench_configs.py --cases Cdev --reps 5 2>/dev/null | grep C-device >> gpurun_out/r2${TAG}_config_c_batch_sweep.jsonl
done
python - <<P
import json
for l in open("gpurun_out/r2${TAG}_config_c_batch_sweep.jsonl"):
    d = json.loads(l); print(d["dtype"], d["batch"], round(d["ms"] * 1e3, 2), "us", round(d["gevals_per_s"]), "Geval/s", d.get("fp64_frac_nominal"))
P
