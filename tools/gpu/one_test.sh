#!/bin/bash
# one GPU test (pytest -k expression).  usage: tools/gpu/one_test.sh EXPR [TAG]
EXPR=${1:?pytest -k expression}; TAG=${2:-one}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "$EXPR" > gpurun_out/r2${TAG}_gputest.log 2>&1; tail -30 gpurun_out/r2${TAG}_gputest.log
