#!/bin/bash
# GPU call o: bucket pipeline variants (FTS_B200_BUCKET_FLAGS: 1 natural task order, 2 no input prefetch)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "config_b or bucket or deep_level" > gpurun_out/r2o_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_gputest.log
tail -3 gpurun_out/r2o_gputest.log
B="python bench.py --permute --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link --no-permuted"
for fl in 0 4; do
FTS_B200_BUCKET_FLAGS=$fl timeout 600 $B > gpurun_out/r2o_bench_f$fl.json 2> gpurun_out/r2o_bench_f$fl.err
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2o_bench_f$fl.json") if l.startswith("{")][-1])
print("flags $fl", d["ms_per_step"], d["step_ms_min"], d["roofline"]["frac"])
P
done
B="python bench.py --permute --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_adaptive1d|k_bucket" -c 8 --csv --log-file gpurun_out/r2o_launches.csv $B > gpurun_out/r2o_ncu.log 2>&1
grep -v "^==" gpurun_out/r2o_launches.csv | cut -d, -f5,15- | head -10
