#!/bin/bash
# GPU call t: stop test fused into the level kernel (one launch per level, one-CTA levels in one launch): tests + 1-GPU sharded3d record, A/B
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r2t_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_gputest.log
tail -15 gpurun_out/r2t_gputest.log
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-host-link --no-permuted"
timeout 600 $B > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
FTS_B200_STEP_FUSED=0 timeout 600 $B > gpurun_out/r2t_bench_twolaunch.json 2> gpurun_out/r2t_bench_twolaunch.err
timeout 300 python tools/bench_configs.py --cases D,Cdev,Edev --reps 20 > gpurun_out/r2t_configs.jsonl 2> gpurun_out/r2t_configs.err
python - <<'P'
import json
for f in ("r2t_bench", "r2t_bench_twolaunch"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        for r in d["sharded3d"]:
            print(f, r["family"], r["level"], r["single"]["ms"], r["single"].get("level_us_rank0"), r["rel_err_vs_closed_form"])
    except Exception as e:
        print(f, "ERR", e)
P
cut -c1-400 gpurun_out/r2t_configs.jsonl
