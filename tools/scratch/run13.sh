(timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider) > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log; grep -E "^FAILED|^E  " gpurun_out/pytest.log | cut -c1-300 | head
timeout 300 python tools/scratch/ord_time2.py 2>&1 | grep ORD
FTS_B200_ORDERED=0 timeout 600 python tools/scratch/ord_time2.py 2>&1 | grep ORD
