(timeout 300 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "pinned or config_b or tail or quad_sem or nvrtc or complement_1d") > gpurun_out/pytest.log 2>&1; tail -1 gpurun_out/pytest.log
grep -q " failed" gpurun_out/pytest.log && { grep -E "^FAILED|Error|error" gpurun_out/pytest.log | cut -c1-300 | head -20; exit 1; }
for r in 1 2; do
for z in 7 3; do echo "ZC=$z"; FTS_B200_ZEROCOPY=$z timeout 120 python tools/e2e_probe.py 2>&1 | tail -1 | cut -c1-190; done
echo "ZC=7 SHIFT=16"; FTS_B200_STREAM_SHIFT=16 timeout 120 python tools/e2e_probe.py 2>&1 | tail -1 | cut -c1-190
echo "ZC=7 SHIFT=18"; FTS_B200_STREAM_SHIFT=18 timeout 120 python tools/e2e_probe.py 2>&1 | tail -1 | cut -c1-190
done
