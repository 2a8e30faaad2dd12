#!/bin/bash
# GPU call p: bucket pipeline: static share of the tasks (flags>>3 eighths; 4 = none) with and without PDL
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --permute --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link --no-permuted"
for cfg in "4 1" "0 1" "16 1" "32 1" "56 1" "4 0" "0 0"; do
set -- $cfg
FTS_B200_BUCKET_FLAGS=$1 FTS_B200_PDL=$2 timeout 600 $B > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2p_bench.json") if l.startswith("{")][-1])
print("flags $1 pdl $2", d["ms_per_step"], d["step_ms_min"], d["roofline"]["frac"])
P
done
B="python bench.py --permute --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link"
FTS_B200_BUCKET_FLAGS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_adaptive1d|k_bucket" -c 8 --csv --log-file gpurun_out/r2p_launches.csv $B > gpurun_out/r2p_ncu.log 2>&1
grep -v "^==" gpurun_out/r2p_launches.csv | cut -d, -f5,15- | head -6
