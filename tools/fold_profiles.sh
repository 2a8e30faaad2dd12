#!/bin/bash
# Fold the captures of one tools/gpu/profile_all.sh call into profiles/ (kernel_profiles.json, ncu summaries, bench lines, launch list).
# usage: tools/fold_profiles.sh TAG   (reads gpurun_out/r2TAG_*)
set -e
TAG=$1
cd "$(dirname "$0")/.."
G=gpurun_out/r2${TAG}
python tools/ncu_profile_json.py ${G}_headline.ncu-rep --kernel-regex k_adaptive1d_tpi --evals 243758272 \
  --files fts_dd.cuh,fts_math.cuh,fts_exp_table.inc,fts_families.cuh,fts_kernels.cuh --note "bench.py --steps 2 --warmup 3 (4th launch), round 2 call $TAG"
python tools/ncu_profile_json.py ${G}_configc.ncu-rep --kernel-regex k_adaptive2d_warp --evals 17305600 \
  --files fts_dd.cuh,fts_math.cuh,fts_families.cuh,fts_kernels.cuh,fts_kernels_cta.cuh --note "config C (4096 boxes), warp-per-integral kernel, round 2 call $TAG"
python tools/ncu_profile_json.py ${G}_log3d.ncu-rep --kernel-regex k_level_grid --evals 941884928 \
  --files fts_dd.cuh,fts_math.cuh,fts_log_table.inc,fts_families.cuh,fts_kernels.cuh,fts_kernels_cta.cuh --note "level 8 of log(1-x)log(1-y)log(1-z), round 2 call $TAG"
python tools/ncu_profile_json.py ${G}_dd_grp.ncu-rep --kernel-regex k_adaptive1d_grp --evals 16646144 \
  --files fts_dd.cuh,fts_dd_tables.inc,fts_math.cuh,fts_families.cuh,fts_kernels.cuh --note "config E (65 536 Double64 quad_cmpl integrals), 4 lanes per integral, round 2 call $TAG"
python tools/ncu_summary.py ${G}_headline.ncu-rep profiles/r2_adaptive1d_tpi_headline_ncu_full.md "headline kernel k_adaptive1d_tpi<double,F1Gauss,0> (config B, 2^20 integrals), round 2 call $TAG"
python tools/ncu_summary.py ${G}_configc.ncu-rep profiles/r2_config_c_warp_kernel_ncu_full.md "config C (4096 boxes, FP64): k_adaptive2d_warp + queue pass, round 2 call $TAG"
python tools/ncu_summary.py ${G}_log3d.ncu-rep profiles/r2_level_grid_log3d_table_log_ncu_full.md "k_level_grid<double,F3Log,3> level 8 (table-driven log), round 2 call $TAG"
python tools/ncu_summary.py ${G}_dd_grp.ncu-rep profiles/r2_adaptive1d_grp_dd_ncu_full.md "config E: k_adaptive1d_grp<dd,C1InvSqrt> (Double64, 4 lanes per integral), round 2 call $TAG"
if [ -f ${G}_bucket.ncu-rep ]; then
  python tools/ncu_summary.py ${G}_bucket.ncu-rep profiles/r2_adaptive1d_bucket_ncu_full.md "config B with shuffled parameters: k_bucket_sort_all + k_adaptive1d_bucket, round 2 call $TAG"
  cp ${G}_launches_permuted.csv profiles/r2${TAG}_bench_launches_permuted.csv
fi
cp ${G}_bench.json profiles/r2${TAG}_bench_1gpu.json
cp ${G}_bench_ref.json profiles/r2${TAG}_bench_ref_1gpu.json
cp ${G}_configs.jsonl profiles/r2${TAG}_configs_c_d_e_nd.jsonl
cp ${G}_launches.csv profiles/r2${TAG}_bench_launches.csv
