#!/bin/bash
# GPU call s: feedback-armed bucket passes: test + bench distribution
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -k "bucket or config_b" > gpurun_out/r2s_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_gputest.log
tail -5 gpurun_out/r2s_gputest.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link"
for m in 1 2; do
FTS_B200_BUCKET=$m timeout 600 $B > gpurun_out/r2s_bench_m$m.json 2> gpurun_out/r2s_bench_m$m.err
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2s_bench_m$m.json") if l.startswith("{")][-1])
print("mode $m", d["ms_per_step"], d["step_ms_median"], d["roofline"]["frac"], d["permuted"], d["e2e"]["ms_per_step"], d["gpu_launches"])
P
done
