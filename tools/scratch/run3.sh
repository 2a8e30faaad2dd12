(timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "pinned or config_b or tail or quad_sem or nvrtc") > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
(FTS_B200_TPI_PERSIST=2 timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "pinned or config_b or tail or quad_sem or nvrtc") > gpurun_out/pytest_p2.log 2>&1; tail -1 gpurun_out/pytest_p2.log
P='import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print("BENCH",sys.argv[1],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["ms_per_step"],d["e2e"]["step_ms_median"])'
for r in 1 2; do
for PS in 0 2; do
  FTS_B200_TPI_PERSIST=$PS timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_p$PS.log 2>&1; python -c "$P" gpurun_out/bench_p$PS.log
  echo "PERSIST=$PS ZC=3"; FTS_B200_TPI_PERSIST=$PS FTS_B200_ZEROCOPY=3 python tools/e2e_probe.py 2>&1 | tail -1 | cut -c1-190
done
done
