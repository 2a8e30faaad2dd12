TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581"
for L in 8 9; do
  timeout 240 $TR tools/run_sharded.py --levels $L --integrand log_3d --reps 5 --exchange peer 2>&1 | grep -E '"case"|rror' | cut -c1-420
done
timeout 240 $TR tools/run_sharded.py --levels 9 --integrand log_3d --reps 5 --exchange nccl 2>&1 | grep -E '"case"|rror' | cut -c1-420
timeout 300 $TR bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | grep -E '"metric"|rror' | cut -c1-4000
FTS_B200_ZEROCOPY=0 timeout 300 $TR bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | grep -E '"metric"|rror' | sed 's/^/ZC0 /' | cut -c1-4000
