// fts_inst.cuh — explicit instantiation + registration of the kernel set of one builtin family.
#pragma once
#include "fts_families.cuh"
#include "fts_internal.h"
#include "fts_kernels.cuh"

namespace fts_host {
struct Registrar {
    Registrar(const KernelEntry* e, int n) { register_kernels(e, n); }
};
}  // namespace fts_host

#define FTS_KPTR(...) ((const void*)(&__VA_ARGS__))

// f32 + f64, reference order and FAST, thread-per-integral kernels of a family
#define FTS_REGISTER_TPI(FAM, FUNCTOR, FIXED_K, ADAPT_K)                                                     \
    static const fts_host::KernelEntry k_entries_##FUNCTOR[] = {                                            \
        {FAM, FTS_F32, fts_host::KID_FIXED_TPI, 0, FTS_KPTR(fts::FIXED_K<float, fts::FUNCTOR, false>)},     \
        {FAM, FTS_F32, fts_host::KID_FIXED_TPI, 1, FTS_KPTR(fts::FIXED_K<float, fts::FUNCTOR, true>)},      \
        {FAM, FTS_F64, fts_host::KID_FIXED_TPI, 0, FTS_KPTR(fts::FIXED_K<double, fts::FUNCTOR, false>)},    \
        {FAM, FTS_F64, fts_host::KID_FIXED_TPI, 1, FTS_KPTR(fts::FIXED_K<double, fts::FUNCTOR, true>)},     \
        {FAM, FTS_F32, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::ADAPT_K<float, fts::FUNCTOR, false>)},     \
        {FAM, FTS_F32, fts_host::KID_ADAPT_TPI, 1, FTS_KPTR(fts::ADAPT_K<float, fts::FUNCTOR, true>)},      \
        {FAM, FTS_F64, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::ADAPT_K<double, fts::FUNCTOR, false>)},    \
        {FAM, FTS_F64, fts_host::KID_ADAPT_TPI, 1, FTS_KPTR(fts::ADAPT_K<double, fts::FUNCTOR, true>)},     \
    };                                                                                                       \
    static fts_host::Registrar k_reg_##FUNCTOR(k_entries_##FUNCTOR, 8);

// pointwise evaluation of a 1D family (fts_eval_batch), f32 + f64
#define FTS_REGISTER_EVAL(FAM, FUNCTOR)                                                                  \
    static const fts_host::KernelEntry k_eval_entries_##FUNCTOR[] = {                                   \
        {FAM, FTS_F32, fts_host::KID_EVAL, 1, FTS_KPTR(fts::k_eval1d<float, fts::FUNCTOR>)},            \
        {FAM, FTS_F64, fts_host::KID_EVAL, 1, FTS_KPTR(fts::k_eval1d<double, fts::FUNCTOR>)},           \
    };                                                                                                   \
    static fts_host::Registrar k_eval_reg_##FUNCTOR(k_eval_entries_##FUNCTOR, 2);

// integrate1D_cmpl of a complement family, f32 + f64, reference order and FAST
#define FTS_REGISTER_FIXED_CMPL(FAM, FUNCTOR)                                                                     \
    static const fts_host::KernelEntry k_fc_entries_##FUNCTOR[] = {                                              \
        {FAM, FTS_F32, fts_host::KID_FIXED_CMPL, 0, FTS_KPTR(fts::k_fixed1d_cmpl_tpi<float, fts::FUNCTOR, false>)},  \
        {FAM, FTS_F32, fts_host::KID_FIXED_CMPL, 1, FTS_KPTR(fts::k_fixed1d_cmpl_tpi<float, fts::FUNCTOR, true>)},   \
        {FAM, FTS_F64, fts_host::KID_FIXED_CMPL, 0, FTS_KPTR(fts::k_fixed1d_cmpl_tpi<double, fts::FUNCTOR, false>)}, \
        {FAM, FTS_F64, fts_host::KID_FIXED_CMPL, 1, FTS_KPTR(fts::k_fixed1d_cmpl_tpi<double, fts::FUNCTOR, true>)},  \
    };                                                                                                            \
    static fts_host::Registrar k_fc_reg_##FUNCTOR(k_fc_entries_##FUNCTOR, 4);

// deep-level tail kernels of the 1D thread-per-integral path (CMPL selects the complement form)
#define FTS_REGISTER_TAIL(FAM, FUNCTOR, CMPL)                                                                        \
    static const fts_host::KernelEntry k_tail_entries_##FUNCTOR[] = {                                               \
        {FAM, FTS_F32, fts_host::KID_ADAPT_TAIL, 0, FTS_KPTR(fts::k_adaptive1d_tail<float, fts::FUNCTOR, false, CMPL>)},   \
        {FAM, FTS_F32, fts_host::KID_ADAPT_TAIL, 1, FTS_KPTR(fts::k_adaptive1d_tail<float, fts::FUNCTOR, true, CMPL>)},    \
        {FAM, FTS_F64, fts_host::KID_ADAPT_TAIL, 0, FTS_KPTR(fts::k_adaptive1d_tail<double, fts::FUNCTOR, false, CMPL>)},  \
        {FAM, FTS_F64, fts_host::KID_ADAPT_TAIL, 1, FTS_KPTR(fts::k_adaptive1d_tail<double, fts::FUNCTOR, true, CMPL>)},   \
    };                                                                                                               \
    static fts_host::Registrar k_tail_reg_##FUNCTOR(k_tail_entries_##FUNCTOR, 4);

// bucket pass of the exp-table families (Float64): arbitrary parameter order in large 1D batches
#define FTS_REGISTER_BUCKET(FAM, FUNCTOR)                                                                            \
    static const fts_host::KernelEntry k_bucket_entries_##FUNCTOR[] = {                                             \
        {FAM, FTS_F64, fts_host::KID_ADAPT_BUCKET, 0, FTS_KPTR(fts::k_adaptive1d_bucket<double, fts::FUNCTOR, false>)},    \
        {FAM, FTS_F64, fts_host::KID_ADAPT_BUCKET, 1, FTS_KPTR(fts::k_adaptive1d_bucket<double, fts::FUNCTOR, true>)},     \
        {FAM, FTS_F64, fts_host::KID_BUCKET_SORT_ALL, 0, FTS_KPTR(fts::k_bucket_sort_all<double, fts::FUNCTOR>)},          \
    };                                                                                                               \
    static fts_host::Registrar k_bucket_reg_##FUNCTOR(k_bucket_entries_##FUNCTOR, 3);

// reference-order cooperative kernels of a 2D / 3D family (CTA per integral, ordered sum), f32 + f64
#define FTS_REGISTER_ORD(FAM, FUNCTOR, ORD_K)                                                            \
    static const fts_host::KernelEntry k_ord_entries_##FUNCTOR[] = {                                    \
        {FAM, FTS_F32, fts_host::KID_ADAPT_ORD, 0, FTS_KPTR(fts::ORD_K<float, fts::FUNCTOR>)},          \
        {FAM, FTS_F64, fts_host::KID_ADAPT_ORD, 0, FTS_KPTR(fts::ORD_K<double, fts::FUNCTOR>)},         \
    };                                                                                                   \
    static fts_host::Registrar k_ord_reg_##FUNCTOR(k_ord_entries_##FUNCTOR, 2);

#define FTS_REGISTER_FIXED_ORD(FAM, FUNCTOR, ORD_K)                                                      \
    static const fts_host::KernelEntry k_ford_entries_##FUNCTOR[] = {                                   \
        {FAM, FTS_F32, fts_host::KID_FIXED_ORD, 0, FTS_KPTR(fts::ORD_K<float, fts::FUNCTOR>)},          \
        {FAM, FTS_F64, fts_host::KID_FIXED_ORD, 0, FTS_KPTR(fts::ORD_K<double, fts::FUNCTOR>)},         \
    };                                                                                                   \
    static fts_host::Registrar k_ford_reg_##FUNCTOR(k_ford_entries_##FUNCTOR, 2);

// adaptive-only families (complement kinds have no exported fixed-grid entry point)
#define FTS_REGISTER_ADAPT_ONLY(FAM, FUNCTOR, ...)                                                           \
    static const fts_host::KernelEntry k_entries_##FUNCTOR[] = {                                            \
        {FAM, FTS_F32, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::__VA_ARGS__<float, fts::FUNCTOR, false>)}, \
        {FAM, FTS_F32, fts_host::KID_ADAPT_TPI, 1, FTS_KPTR(fts::__VA_ARGS__<float, fts::FUNCTOR, true>)},  \
        {FAM, FTS_F64, fts_host::KID_ADAPT_TPI, 0, FTS_KPTR(fts::__VA_ARGS__<double, fts::FUNCTOR, false>)},\
        {FAM, FTS_F64, fts_host::KID_ADAPT_TPI, 1, FTS_KPTR(fts::__VA_ARGS__<double, fts::FUNCTOR, true>)}, \
    };                                                                                                       \
    static fts_host::Registrar k_reg_##FUNCTOR(k_entries_##FUNCTOR, 4);


#include "fts_kernels_cta.cuh"

// cooperative FAST kernels (CTA per integral) of a family, f32 + f64
#define FTS_REGISTER_CTA(FAM, FUNCTOR, FIXED_K, ADAPT_K)                                                 \
    static const fts_host::KernelEntry k_cta_entries_##FUNCTOR[] = {                                    \
        {FAM, FTS_F32, fts_host::KID_FIXED_CTA, 1, FTS_KPTR(fts::FIXED_K<float, fts::FUNCTOR>)},        \
        {FAM, FTS_F64, fts_host::KID_FIXED_CTA, 1, FTS_KPTR(fts::FIXED_K<double, fts::FUNCTOR>)},       \
        {FAM, FTS_F32, fts_host::KID_ADAPT_CTA, 1, FTS_KPTR(fts::ADAPT_K<float, fts::FUNCTOR>)},        \
        {FAM, FTS_F64, fts_host::KID_ADAPT_CTA, 1, FTS_KPTR(fts::ADAPT_K<double, fts::FUNCTOR>)},       \
    };                                                                                                   \
    static fts_host::Registrar k_cta_reg_##FUNCTOR(k_cta_entries_##FUNCTOR, 4);

#define FTS_REGISTER_CTA_ADAPT_ONLY(FAM, FUNCTOR, ADAPT_K)                                               \
    static const fts_host::KernelEntry k_cta_entries_##FUNCTOR[] = {                                    \
        {FAM, FTS_F32, fts_host::KID_ADAPT_CTA, 1, FTS_KPTR(fts::ADAPT_K<float, fts::FUNCTOR>)},        \
        {FAM, FTS_F64, fts_host::KID_ADAPT_CTA, 1, FTS_KPTR(fts::ADAPT_K<double, fts::FUNCTOR>)},       \
    };                                                                                                   \
    static fts_host::Registrar k_cta_reg_##FUNCTOR(k_cta_entries_##FUNCTOR, 2);

// warp-per-integral kernel of a 2D family (shallow levels of large batches), f32 + f64
#define FTS_REGISTER_WARP2D(FAM, FUNCTOR)                                                                \
    static const fts_host::KernelEntry k_w2_entries_##FUNCTOR[] = {                                     \
        {FAM, FTS_F32, fts_host::KID_ADAPT_WARP, 1, FTS_KPTR(fts::k_adaptive2d_warp<float, fts::FUNCTOR>)},  \
        {FAM, FTS_F64, fts_host::KID_ADAPT_WARP, 1, FTS_KPTR(fts::k_adaptive2d_warp<double, fts::FUNCTOR>)}, \
    };                                                                                                   \
    static fts_host::Registrar k_w2_reg_##FUNCTOR(k_w2_entries_##FUNCTOR, 2);

// one-level whole-grid kernels (single large integral / multi-GPU shards), f32 + f64
#define FTS_REGISTER_GRID(FAM, FUNCTOR, D)                                                               \
    static const fts_host::KernelEntry k_grid_entries_##FUNCTOR[] = {                                   \
        {FAM, FTS_F32, fts_host::KID_LEVEL_GRID, 1, FTS_KPTR(fts::k_level_grid<float, fts::FUNCTOR, D>)},  \
        {FAM, FTS_F64, fts_host::KID_LEVEL_GRID, 1, FTS_KPTR(fts::k_level_grid<double, fts::FUNCTOR, D>)}, \
    };                                                                                                   \
    static fts_host::Registrar k_grid_reg_##FUNCTOR(k_grid_entries_##FUNCTOR, 2);
