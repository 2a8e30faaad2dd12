(timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider) > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
grep -q " failed" gpurun_out/pytest.log && { grep -E "^FAILED" gpurun_out/pytest.log | cut -c1-200; exit 1; }
P='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print("BENCH",sys.argv[1],d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"])'
for B in 256 128 64; do
  FTS_B200_TPI_BLOCK=$B timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b$B.log 2>&1; python -c "$P" gpurun_out/bench_b$B.log
done
(timeout 300 python tools/bench_configs.py --reps 5 --cases Adev,Dk) > gpurun_out/configs_x.log 2>&1; grep -E "A-device|exp_3d" gpurun_out/configs_x.log | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pre_ncu.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adaptive1d_tpi -c 1 -s 4 -o gpurun_out/r1_tpi_v4 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v4.log 2>&1
ls -la gpurun_out/*.ncu-rep
