TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
for ex in peer nccl; do for L in 8 9; do
  timeout 240 $TR tools/run_sharded.py --levels $L --integrand log_3d --reps 5 --exchange $ex 2>&1 | grep -E '"case"|rror' | sed "s/^/EX=$ex /" | cut -c1-420
done; done
timeout 240 $TR tools/run_sharded.py --levels 9 --integrand exp_3d --reps 5 --exchange peer 2>&1 | grep -E '"case"|rror' | sed "s/^/EX=peer /" | cut -c1-420
timeout 300 $TR bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | grep -E '"metric"|rror' | cut -c1-3000
