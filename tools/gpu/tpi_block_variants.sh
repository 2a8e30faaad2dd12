#!/bin/bash
# threads per CTA of the 1D thread-per-integral kernels (FTS_B200_TPI_BLOCK): bench.py headline + e2e per value.  usage: tools/gpu/tpi_block_variants.sh TAG
TAG=${1:-blk}
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link --no-other-configs --no-permuted"
for v in ${BLOCKS:-128 256 192 160 96 128}; do
  FTS_B200_TPI_BLOCK=$v timeout 300 $B > gpurun_out/r2${TAG}_bench_b$v.json 2> gpurun_out/r2${TAG}_bench_b$v.err
  python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench_b$v.json") if l.startswith("{")][-1])
print("block $v", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"])
P
done
