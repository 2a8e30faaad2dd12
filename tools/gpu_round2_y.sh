#!/bin/bash
# GPU call y: table-driven Double64 exp/log + 4 lanes per integral for Double64 1D batches (config E)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_bench_contract.py::test_our_arm_line > gpurun_out/r2y_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_gputest.log
tail -8 gpurun_out/r2y_gputest.log
for G in 4 2 1; do
FTS_B200_DD_GROUP=$G timeout 300 python tools/bench_configs.py --cases Edev,E --reps 20 > gpurun_out/r2y_configs_g$G.jsonl 2> gpurun_out/r2y_configs_g$G.err
cut -c1-330 gpurun_out/r2y_configs_g$G.jsonl
done
