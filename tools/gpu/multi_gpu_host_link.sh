#!/bin/bash
# multi-GPU call: bench.py both arms under torchrun, host-link probe, sharded 3D integral.   usage: tools/gpu/multi_gpu_host_link.sh N TAG
N=${1:-2}; TAG=${2:-m}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2${TAG}_topo.txt 2>&1; lscpu | grep -i -E "numa|model name|^cpu\(s\)|socket|thread" >> gpurun_out/r2${TAG}_topo.txt; numactl -H >> gpurun_out/r2${TAG}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2${TAG}_bench_${N}gpu.json 2> gpurun_out/r2${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 2 > gpurun_out/r2${TAG}_bench_ref_${N}gpu.json 2> gpurun_out/r2${TAG}_bench_ref_${N}gpu.err; echo "ref rc=$?"
timeout 600 $TR tools/host_link_probe.py --reps 100 > gpurun_out/r2${TAG}_hostlink_${N}gpu.jsonl 2> gpurun_out/r2${TAG}_hostlink_${N}gpu.err; echo "probe rc=$?"
tail -c 1500 gpurun_out/r2${TAG}_bench_${N}gpu.json; tail -3 gpurun_out/r2${TAG}_bench_${N}gpu.err; cat gpurun_out/r2${TAG}_hostlink_${N}gpu.jsonl | cut -c1-400; tail -5 gpurun_out/r2${TAG}_hostlink_${N}gpu.err
