(timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "pinned or config_b") > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
for r in 1 2 3; do
for w in 0 1 2; do echo "ZC=3 PF=$w"; FTS_B200_ZEROCOPY=3 FTS_B200_PREFETCH_WAVES=$w timeout 120 python tools/e2e_probe.py 2>&1 | tail -1 | cut -c1-190; done
done
