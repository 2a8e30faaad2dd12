#!/bin/bash
# GPU call z: dd_rsqrt + one-reciprocal dd division (config E invsqrt family), empty 2D queue pass returns before its table set-up (config C)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_bench_contract.py::test_our_arm_line > gpurun_out/r2z_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_gputest.log
tail -4 gpurun_out/r2z_gputest.log
for G in 4 1; do
FTS_B200_DD_GROUP=$G timeout 300 python tools/bench_configs.py --cases Edev,Cdev --reps 20 > gpurun_out/r2z_configs_g$G.jsonl 2> gpurun_out/r2z_configs_g$G.err
cut -c1-330 gpurun_out/r2z_configs_g$G.jsonl
done
