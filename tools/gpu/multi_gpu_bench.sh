#!/bin/bash
# multi-GPU check of the final state: bench.py both arms under torchrun + smoke.   usage: tools/gpu/multi_gpu_bench.sh N TAG
N=${1:-2}; TAG=${2:-m}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2${TAG}_bench_${N}gpu.json 2> gpurun_out/r2${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 2 > gpurun_out/r2${TAG}_bench_ref_${N}gpu.json 2> gpurun_out/r2${TAG}_bench_ref_${N}gpu.err; echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2${TAG}_smoke.log
python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_host_link"), d["permuted"]["frac"])
for x in d["sharded3d"]: print(x["family"], x["level"], x["single"]["ms"], x.get("fused_peer", {}).get("ms"), x.get("fused_peer", {}).get("speedup_vs_single"), x.get("nccl_allgather", {}).get("ms"))
r = json.loads([l for l in open("gpurun_out/r2${TAG}_bench_ref_${N}gpu.json") if l.startswith("{")][-1])
print("ref", r["value"], r["cpu_baseline"]["cores"])
P
