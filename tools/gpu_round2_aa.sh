#!/bin/bash
# GPU call aa: Double64 2D/3D thread-per-integral kernels + quad_split in Double64
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "generic_nd or double64_2d or nvrtc" > gpurun_out/r2aa_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_gputest.log
tail -30 gpurun_out/r2aa_gputest.log
