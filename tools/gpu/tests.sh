#!/bin/bash
# full GPU suite only.  usage: tools/gpu/tests.sh TAG
TAG=${1:-t1}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2${TAG}_gputest.log
tail -12 gpurun_out/r2${TAG}_gputest.log
