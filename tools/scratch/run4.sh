(timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider) > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
grep -q " failed" gpurun_out/pytest.log && { grep -E "^FAILED|Error" gpurun_out/pytest.log | cut -c1-300 | head -20; exit 1; }
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
(timeout 300 python tools/bench_configs.py --reps 5 --cases Apin,C) > gpurun_out/configs_a.log 2>&1; grep -E "\"case\"" gpurun_out/configs_a.log | cut -c1-300
