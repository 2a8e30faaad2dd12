#!/bin/bash
# final bench lines on one GPU: bench.py both arms + tools/bench_configs.py.  usage: tools/gpu/bench_lines.sh TAG
TAG=${1:-b1}
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r2${TAG}_bench.json 2> gpurun_out/r2${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2${TAG}_bench_ref.json 2> gpurun_out/r2${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/bench_configs.py --cases Cdev,Dk,Edev,D,ND --reps 10 > gpurun_out/r2${TAG}_configs.jsonl 2> gpurun_out/r2${TAG}_configs.err
timeout 600 python bench.py > gpurun_out/r2${TAG}_bench_default.json 2> gpurun_out/r2${TAG}_bench_default.err; echo "default rc=$?"
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_executed"), d["roofline"].get("frac_pipe_slots"), d["roofline"]["traffic"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["cold"]["ms_per_step"], "perm", d["permuted"]["ms_per_step"], d["permuted"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["value_avx"])
print([(x["case"], x.get("dtype") or x.get("family"), round(x["ms"], 4)) for x in d["other_configs"]])
for x in d["sharded3d"]: print(x["family"], x["level"], x["single"]["ms"])
P
cat gpurun_out/r2${TAG}_configs.jsonl | cut -c1-330
