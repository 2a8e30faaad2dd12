#!/bin/bash
# config C (4096 2D complement boxes) device-timed over alternative builds of the library + the 2D tests on each.
# usage: tools/gpu/config_c_variants.sh TAG v1 v2 ...   ("base" = lib/, NAME = lib_NAME/)
TAG=$1; shift
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  L=$PWD/fasttanhsinhquadrature.jl_b200/lib_$v/libfts_b200.so
  [ $v = base ] && L=$PWD/fasttanhsinhquadrature.jl_b200/lib/libfts_b200.so
  FTS_B200_LIB=$L timeout 300 python tools/bench_configs.py --cases Cdev --reps 10 > gpurun_out/r2${TAG}_configs_$v.jsonl 2> gpurun_out/r2${TAG}_configs_$v.err
  python - <<P
import json
for l in open("gpurun_out/r2${TAG}_configs_$v.jsonl"):
    d = json.loads(l)
    if d.get("case") == "C-device": print("$v", d["dtype"], "graph", round(d["ms_graph_replay"] * 1e3, 2), "us  loop", round(d["ms_python_loop"] * 1e3, 2), "us", d.get("fp64_frac_nominal"), d.get("max_rel_err"))
P
done
done
for v in "$@"; do
  [ $v = base ] && continue
  FTS_B200_LIB=$PWD/fasttanhsinhquadrature.jl_b200/lib_$v/libfts_b200.so timeout 600 python -m pytest tests -m gpu -q --maxfail=5 -k "2d or 2D or config_c or sweep or kat or cmpl or complement or edge" > gpurun_out/r2${TAG}_gputest_$v.log 2>&1
  tail -3 gpurun_out/r2${TAG}_gputest_$v.log
done
