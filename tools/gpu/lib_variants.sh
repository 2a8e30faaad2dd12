#!/bin/bash
# headline bench (device-resident + e2e) over alternative builds of the library.  usage: tools/gpu/lib_variants.sh TAG v1 v2 ...
# (a variant NAME is fasttanhsinhquadrature.jl_b200/lib_NAME/libfts_b200.so; "base" = lib/)
TAG=$1; shift
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-permuted --no-host-link --no-other-configs"
for rep in 1 2; do
for v in "$@"; do
  L=$PWD/fasttanhsinhquadrature.jl_b200/lib_$v/libfts_b200.so
  [ $v = base ] && L=$PWD/fasttanhsinhquadrature.jl_b200/lib/libfts_b200.so
  FTS_B200_LIB=$L timeout 300 $B > gpurun_out/r2${TAG}_bench_$v.json 2> gpurun_out/r2${TAG}_bench_$v.err
  python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench_$v.json") if l.startswith("{")][-1])
print("$v", round(d["ms_per_step"], 5), round(d["step_ms_min"], 5), round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["ms_per_step"], 5), d["converged"])
P
done
done
