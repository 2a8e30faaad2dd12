#!/bin/bash
# GPU call r: bucket pipeline final shape: full GPU tests, bench in both orders
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r2r_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_gputest.log
tail -5 gpurun_out/r2r_gputest.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link"
timeout 600 $B > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2r_bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["roofline"]["frac"], d["permuted"], d["e2e"]["ms_per_step"])
P
