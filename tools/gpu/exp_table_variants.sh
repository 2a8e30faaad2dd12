#!/bin/bash
# exp-table size experiment: bench.py (headline + permuted + e2e) with the default (128-entry) and FTS_EXP_BITS=9 builds; ulp test on the variant
TAG=${1:-e1}
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-sharded3d --no-host-link --no-other-configs"
for v in 7 9 7 9; do
  L=fasttanhsinhquadrature.jl_b200/lib_e$v/libfts_b200.so
  [ $v = 7 ] && L=fasttanhsinhquadrature.jl_b200/lib/libfts_b200.so
  FTS_B200_LIB=$PWD/$L timeout 300 $B > gpurun_out/r2${TAG}_bench_e$v.json 2> gpurun_out/r2${TAG}_bench_e$v.err
  python - <<P
import json
d = json.loads([l for l in open("gpurun_out/r2${TAG}_bench_e$v.json") if l.startswith("{")][-1])
print("bits $v", d["ms_per_step"], d["roofline"]["frac"], {k: d["permuted"].get(k) for k in ("ms_per_step", "frac")}, d["e2e"]["ms_per_step"])
P
done
FTS_B200_LIB=$PWD/fasttanhsinhquadrature.jl_b200/lib_e9/libfts_b200.so timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "math or config_b or bucket or golden or csv or kat or sweep" > gpurun_out/r2${TAG}_gputest_e9.log 2>&1; tail -15 gpurun_out/r2${TAG}_gputest_e9.log
