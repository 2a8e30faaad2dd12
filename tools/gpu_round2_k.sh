#!/bin/bash
# GPU call k: bucket pass for arbitrary parameter order (headline kernel), tests + bench both orders
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "config_b or bucket or deep_level or pinned or reentrant or complement_1d" > gpurun_out/r2k_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_gputest.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link"
timeout 600 $B > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
FTS_B200_BUCKET=0 timeout 600 $B > gpurun_out/r2k_bench_nobucket.json 2> gpurun_out/r2k_bench_nobucket.err
tail -8 gpurun_out/r2k_gputest.log
python - <<'P'
import json
for f in ("r2k_bench", "r2k_bench_nobucket"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["roofline"]["frac"], d["permuted"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
P
