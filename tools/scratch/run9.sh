(timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider) > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
grep -q " failed" gpurun_out/pytest.log && { grep -E "^FAILED|^E  " gpurun_out/pytest.log | cut -c1-300 | head -20; }
(FTS_B200_WARP2D_LEVELS=-1 timeout 300 python tools/bench_configs.py --reps 10 --cases Cdev) 2>&1 | grep '"case": "C' | cut -c1-400
for lv in 4 5 6; do (FTS_B200_WARP2D_LEVELS=$lv timeout 300 python tools/bench_configs.py --reps 10 --cases Cdev) 2>&1 | grep '"case": "C' | cut -c1-400; done
