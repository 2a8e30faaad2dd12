// fts_internal.h — host-side internals shared by the translation units of libfts_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/fts_b200.h"

namespace fts_host {

// kernel ids of the registry (one entry per family x dtype x kernel x FAST)
enum KernelId : int {
    KID_FIXED_TPI = 0,     // k_fixed{1,2,3}d_tpi          thread per integral
    KID_ADAPT_TPI = 1,     // k_adaptive{1d,1d_cmpl,2d,3d}_tpi
    KID_FIXED_CTA = 2,     // k_fixed*_cta                 CTA per integral (FAST only)
    KID_ADAPT_CTA = 3,     // k_adaptive*_cta              CTA per integral (FAST only)
    KID_LEVEL_GRID = 4,    // k_level*_grid                many CTAs per integral, one launch per level
    KID_ADAPT_TAIL = 5,    // k_adaptive1d_tail               deep levels of queued 1D integrals (reference order)
    KID_ADAPT_ORD = 6,     // k_adaptive{2,3}d_ord            CTA per integral, REFERENCE summation order (small 2D/3D batches)
    KID_FIXED_ORD = 7,     // k_fixed{2,3}d_ord               CTA per integral, REFERENCE order (small 2D/3D batches)
    KID_EVAL = 8,          // k_eval1d                        pointwise integrand values (1D kinds)
    KID_FIXED_CMPL = 9,    // k_fixed1d_cmpl_tpi              integrate1D_cmpl (fixed grid, complement signature)
    KID_ADAPT_WARP = 10,   // k_adaptive2d_warp               warp per integral, shallow levels (large 2D batches, FAST only)
    KID_ADAPT_BUCKET = 11, // k_adaptive1d_bucket             warp tasks over bound-ordered chunks (1D exp families, Float64, arbitrary parameter order)
    KID_ADAPT_GRP = 12,    // k_adaptive1d_grp                a group of 4 lanes per integral (Double64, 1D kinds); `fast` = 1: groups of 2
    KID_FIXED_ND = 13,     // k_fixed_nd                      generic n-D tensor product, fixed grid (NVRTC integrands of kind FTS_KIND_ND + D)
    KID_BUCKET_SORT_ALL = 14, // k_bucket_sort_all            keys + per-chunk ordering of a whole batch (first kernel of a call in arbitrary parameter order)
    KID_COUNT = 15
};

struct KernelEntry {
    int family, dtype, kid, fast;
    const void* fn;  // __global__ host stub (static) or cudaKernel_t (NVRTC)
};

void register_kernels(const KernelEntry* entries, int count);
const void* find_kernel(int family, int dtype, int kid, int fast);

int set_error(int code, const std::string& msg);
int cuda_error(cudaError_t e, const char* what);

struct FamilyInfo {
    int family, kind, nparams;
    const char* name;
};
const FamilyInfo* family_info(int family);

inline size_t dtype_size(int dtype) { return dtype == FTS_F32 ? 4 : (dtype == FTS_F64 ? 8 : 16); }

}  // namespace fts_host

struct fts_tables_s {
    int dtype, n;
    void* d_grid;         // Node<T>[n]
    unsigned char h[16];  // one element
};

struct fts_cache_s {
    int dtype, D, kind, max_levels;  // D = 0 for 1D caches; kind: 0 regular, 1 complement
    long long count;                 // nodes in the flat level arrays
    void* d_nodes;                   // Node<T>[count]
    void* d_nodes_c;                 // NodeC<T>[count] (complement caches only)
    unsigned char tm[16], ix[32], iw[32], ic[32];
};

struct fts_integrand_s {
    int builtin, family, kind, nparams;
    void* module;  // NVRTC integrands: compiled module (fts_nvrtc.cpp)
};

// fts_nvrtc.cpp: kernel of an NVRTC module (loads the cubin on first use); null + last_error on failure
const void* fts_nvrtc_kernel(void* module, int dtype, int kid, int fast);

#define FTS_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return fts_host::cuda_error(e__, #call);   \
    } while (0)
