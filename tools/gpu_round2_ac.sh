#!/bin/bash
# GPU call ac: detection pass of the bucket path leaves before the table fill; full GPU tests
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_bench_contract.py::test_our_arm_line > gpurun_out/r2ac_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_gputest.log
tail -4 gpurun_out/r2ac_gputest.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link"
for m in 1 2; do
FTS_B200_BUCKET=$m timeout 600 $B > gpurun_out/r2ac_bench_m$m.json 2> gpurun_out/r2ac_bench_m$m.err
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2ac_bench_m$m.json") if l.startswith("{")][-1])
print("mode $m", d["ms_per_step"], d["step_ms_median"], d["roofline"]["frac"], d["permuted"], d["e2e"]["ms_per_step"], d["gpu_launches"])
P
done
