P='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print("BENCH",sys.argv[1],d["ms_per_step"],d["step_ms_min"],d["roofline"]["frac"])'
for r in 1 2; do for B in 256 128; do
  FTS_B200_TPI_BLOCK=$B timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b$B.log 2>&1; python -c "$P" gpurun_out/bench_b$B.log
done; done
for B in 256 128; do echo "BLOCK=$B"; (FTS_B200_TPI_BLOCK=$B timeout 300 python tools/bench_configs.py --reps 10 --cases Adev) 2>&1 | grep "A-device" | cut -c1-150; done
