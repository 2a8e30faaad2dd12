#!/bin/bash
# GPU call u: levels 0..7 of a grid integral in one cooperative launch (generation barrier): tests + 1-GPU sharded3d record, A/B
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_bench_contract.py::test_our_arm_line > gpurun_out/r2w_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_gputest.log
tail -8 gpurun_out/r2w_gputest.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-host-link --no-permuted"
for R in 7 0; do
FTS_B200_GRID_RANGE=$R timeout 600 $B > gpurun_out/r2w_bench_range$R.json 2> gpurun_out/r2w_bench_range$R.err
done
timeout 300 python tools/bench_configs.py --cases D --reps 20 > gpurun_out/r2w_configs.jsonl 2> gpurun_out/r2w_configs.err
FTS_B200_GRID_RANGE=0 timeout 300 python tools/bench_configs.py --cases D --reps 20 >> gpurun_out/r2w_configs.jsonl 2>> gpurun_out/r2w_configs.err
python - <<'P'
import json
for f in ("r2w_bench_range7", "r2w_bench_range0"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        for r in d["sharded3d"]:
            print(f, r["family"], r["level"], r["single"]["ms"], r["single"].get("level_us_rank0"), r["rel_err_vs_closed_form"])
    except Exception as e:
        print(f, "ERR", e)
P
cut -c1-330 gpurun_out/r2w_configs.jsonl
