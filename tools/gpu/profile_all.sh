#!/bin/bash
# GPU profiling call: full tests, the bench line, and `ncu --set full` captures of the kernels recorded in
# profiles/kernel_profiles.json (headline, config C warp kernel, log_3d level kernel) + the Double64 group kernel.
# usage: tools/gpu/profile_all.sh TAG
TAG=${1:-prof}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r2${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2${TAG}_gputest.log
tail -4 gpurun_out/r2${TAG}_gputest.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2${TAG}_bench.json 2> gpurun_out/r2${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2${TAG}_bench_ref.json 2> gpurun_out/r2${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/bench_configs.py --cases Cdev,Dk,Edev,D,ND --reps 10 > gpurun_out/r2${TAG}_configs.jsonl 2> gpurun_out/r2${TAG}_configs.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link --no-other-configs"
timeout 300 $B > gpurun_out/r2${TAG}_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_tpi -s 3 -c 1 -f -o gpurun_out/r2${TAG}_headline $B > gpurun_out/r2${TAG}_ncu_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2${TAG}_launches.csv $B > gpurun_out/r2${TAG}_ncu_l.log 2>&1
C="python tools/bench_configs.py --cases Cdev --reps 1"
timeout 200 $C > gpurun_out/r2${TAG}_plain_c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive2d -s 6 -c 2 -f -o gpurun_out/r2${TAG}_configc $C > gpurun_out/r2${TAG}_ncu_c.log 2>&1
D="python tools/bench_configs.py --cases Dk --reps 1"
FTS_BENCH_DK=log_3d:8 timeout 200 $D > gpurun_out/r2${TAG}_plain_d.log 2>&1 && FTS_BENCH_DK=log_3d:8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_level_grid -s 2 -c 1 -f -o gpurun_out/r2${TAG}_log3d $D > gpurun_out/r2${TAG}_ncu_d.log 2>&1
Ee="python tools/bench_configs.py --cases Edev --reps 1"
timeout 200 $Ee > gpurun_out/r2${TAG}_plain_e.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_grp -s 2 -c 1 -f -o gpurun_out/r2${TAG}_dd_grp $Ee > gpurun_out/r2${TAG}_ncu_e.log 2>&1
P="python bench.py --permute --steps 2 --warmup 3 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link --no-other-configs"
timeout 200 $P > gpurun_out/r2${TAG}_plain_p.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_adaptive1d_bucket|k_bucket_sort_all" -s 4 -c 2 -f -o gpurun_out/r2${TAG}_bucket $P > gpurun_out/r2${TAG}_ncu_p.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_adaptive1d|k_bucket" -c 24 --csv --log-file gpurun_out/r2${TAG}_launches_permuted.csv $P > gpurun_out/r2${TAG}_ncu_pl.log 2>&1
tail -n 2 gpurun_out/r2${TAG}_ncu_b.log gpurun_out/r2${TAG}_ncu_p.log gpurun_out/r2${TAG}_ncu_c.log gpurun_out/r2${TAG}_ncu_d.log gpurun_out/r2${TAG}_ncu_e.log
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["permuted"]["step_ms_median"], d["cpu_baseline"], [ (x["case"], x.get("dtype") or x.get("family"), round(x["ms"],4)) for x in d["other_configs"]])
P
