#!/bin/bash
# sort-all mode of the bucket passes: tests + bench (headline + permuted) per FTS_B200_BUCKET mode.  usage: tools/gpu/bucket_passes.sh TAG
TAG=${1:-p3}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -k "bucket or config_b or deep_level or permut or full_size" > gpurun_out/r2${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2${TAG}_gputest.log
tail -5 gpurun_out/r2${TAG}_gputest.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link --no-other-configs"
for m in 1 2 3 1; do
  FTS_B200_BUCKET=$m timeout 300 $B > gpurun_out/r2${TAG}_bench_m$m.json 2> gpurun_out/r2${TAG}_bench_m$m.err
  python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench_m$m.json") if l.startswith("{")][-1])
print("mode $m", d["ms_per_step"], d["roofline"]["frac"], {k: d["permuted"].get(k) for k in ("step_ms_median", "ms_per_step", "frac")}, d["gpu_launches"])
P
done
B2="python bench.py --permute --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_adaptive1d|k_bucket" -c 12 --csv --log-file gpurun_out/r2${TAG}_launches.csv $B2 > gpurun_out/r2${TAG}_ncu.log 2>&1
grep -v "^==" gpurun_out/r2${TAG}_launches.csv | cut -d, -f5,15- | head -14
