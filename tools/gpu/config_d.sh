#!/bin/bash
# config D (one 2D/3D integral on many CTAs): 3D / sharded tests, whole-integral timings, per-level device timestamps.  usage: tools/gpu/config_d.sh TAG
TAG=${1:-d1}
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "3d or 3D or sharded or config_d or grid or quad or csv" > gpurun_out/r2${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2${TAG}_gputest.log
tail -4 gpurun_out/r2${TAG}_gputest.log
timeout 300 python tools/bench_configs.py --cases D --reps 20 > gpurun_out/r2${TAG}_configs.jsonl 2> gpurun_out/r2${TAG}_configs.err
cut -c1-260 gpurun_out/r2${TAG}_configs.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-permuted --no-host-link --no-other-configs > gpurun_out/r2${TAG}_bench.json 2> gpurun_out/r2${TAG}_bench.err
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench.json") if l.startswith("{")][-1])
for x in d["sharded3d"]: print(x["family"], x["level"], x["single"]["ms"], x["single"].get("level_us_rank0"))
P
