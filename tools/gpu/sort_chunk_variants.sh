#!/bin/bash
# bucket-pass chunk size experiment: bench.py (headline + permuted) with SORT_CHUNK 1024 / 2048 / 4096 builds.  usage: tools/gpu/sort_chunk_variants.sh TAG
TAG=${1:-p1}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link --no-other-configs"
for v in ${CHUNKS:-1024 512 256 1024 512}; do
  L=fasttanhsinhquadrature.jl_b200/lib_c$v/libfts_b200.so
  [ $v = 1024 ] && L=fasttanhsinhquadrature.jl_b200/lib/libfts_b200.so
  FTS_B200_LIB=$PWD/$L timeout 300 $B > gpurun_out/r2${TAG}_bench_c$v.json 2> gpurun_out/r2${TAG}_bench_c$v.err
  python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench_c$v.json") if l.startswith("{")][-1])
print("chunk $v", d["ms_per_step"], d["roofline"]["frac"], {k: d["permuted"].get(k) for k in ("step_ms_median", "ms_per_step", "frac")})
P
done
