#!/bin/bash
# multi-GPU call v: sharded3d record with the one-launch level range vs per-level launches vs two launches per level.  usage: gpu_round2_v.sh N
N=${1:-8}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-host-link --no-permuted"
timeout 300 $TR $B > gpurun_out/r2x_bench_${N}gpu_range7.json 2> gpurun_out/r2x_bench_${N}gpu_range7.err; echo "rc=$?"
FTS_B200_GRID_RANGE=0 timeout 300 $TR $B > gpurun_out/r2x_bench_${N}gpu_range0.json 2> gpurun_out/r2x_bench_${N}gpu_range0.err; echo "rc=$?"
FTS_B200_STEP_FUSED=0 timeout 300 $TR $B > gpurun_out/r2x_bench_${N}gpu_twolaunch.json 2> gpurun_out/r2x_bench_${N}gpu_twolaunch.err; echo "rc=$?"
python - <<P
import json
for f in ("range7", "range0", "twolaunch"):
    try:
        d = json.loads(open(f"gpurun_out/r2x_bench_${N}gpu_{f}.json").read().strip().splitlines()[-1])
        for r in d["sharded3d"]:
            fp = r["fused_peer"]
            print(f, r["family"], r["level"], "single", round(r["single"]["ms"], 4), "fused", round(fp["ms"], 4), "x", round(fp["speedup_vs_single"], 2), "nccl", round(r["nccl_allgather"]["ms"], 4), fp.get("level_us_rank0"), fp["exchange_wait_us_per_level_max_over_ranks"], fp["identical_on_all_ranks"])
    except Exception as e:
        print(f, "ERR", repr(e))
P
tail -3 gpurun_out/r2x_bench_${N}gpu_range7.err
