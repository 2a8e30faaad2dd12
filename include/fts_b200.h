/*
 * fts_b200.h — C ABI of libfts_b200.so, the B200 (sm_100a) implementation of the
 * FastTanhSinhQuadrature.jl hot path: weighted tanh-sinh node sums (fixed grids) and the
 * level-by-level adaptive refinement, batched over many integrals.
 *
 * The reference (pure Julia, /root/reference) has no FFI boundary: the integrand is a Julia
 * closure inlined into the loop.  This header IS the boundary a Julia shim binds with `ccall`
 * (julia/FastTanhSinhQuadratureB200).  Every entry point names the reference
 * function (file:line under /root/reference) it replaces.
 *
 * Conventions
 *  - every function returns an int status (FTS_OK == 0); fts_last_error() has the message
 *    (thread-local).  No C++ exception crosses the ABI.
 *  - scalar arguments of the working precision are passed as `const void*` pointing at ONE
 *    element of the dtype (float / double / fts_dd), arrays as `const void*` to `count` elements.
 *  - `*_batch` entry points take HOST buffers (caller-owned, only read/written during the call)
 *    and move them over PCIe themselves; they are re-entrant — every host thread gets its own
 *    stream and device scratch, two threads run two calls side by side; `*_batch_dev` take DEVICE pointers on `stream`
 *    (a cudaStream_t passed as void*, NULL = legacy default stream) and are asynchronous.
 *  - layout: bounds `low`,`up` are [batch][D] row-major (a Julia Vector{SVector{D,T}});
 *    `bounds_stride` = D for per-integral boxes or 0 to broadcast one box to the whole batch.
 *    `params` is [batch][nparams] row-major (`params_stride` = nparams, or 0 to broadcast).
 *  - outputs: value[batch] (dtype), err[batch] (dtype, nullable), levels[batch] (int32, nullable),
 *    flags[batch] (int32, nullable; bit FTS_FLAG_*).
 *  - handles are opaque, immutable after creation, owned by the library and freed explicitly.
 */
#ifndef FTS_B200_H
#define FTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTS_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------------ */
enum {
    FTS_OK = 0,
    FTS_ERR_INVALID_ARGUMENT = 1,   /* Julia shim maps to ArgumentError (core.jl:41-42, caches.jl:26-27) */
    FTS_ERR_DIMENSION_MISMATCH = 2, /* DimensionMismatch (integrate2D.jl:149-150, quad.jl:70) */
    FTS_ERR_UNSUPPORTED = 3,        /* e.g. ">3D" (quad.jl:88), family/dtype combination not built */
    FTS_ERR_CUDA = 4,               /* any CUDA runtime failure; message carries cudaGetErrorString */
    FTS_ERR_NVRTC = 5,              /* user integrand failed to compile; log returned to caller */
    FTS_ERR_NO_DEVICE = 6,          /* no usable sm_100 device: there is NO CPU fallback */
    FTS_ERR_INTERNAL = 7
};

/* ---- element types (dtype ∈ {f32,f64,dd64}) ------------------------------------ */
typedef enum { FTS_F32 = 0, FTS_F64 = 1, FTS_DD64 = 2 } fts_dtype;
#define FTS_ND_MAX 8

/* Double64 element: unevaluated sum hi+lo, |lo| <= ulp(hi)/2 (DoubleFloats.jl layout). */
typedef struct { double hi, lo; } fts_dd;

/* ---- summation order --------------------------------------------------------------------- */
typedef enum {
    /* exact sequential order of the reference's scalar kernels (what quad/quad_cmpl/quad_split
       and integrate*D / adaptive_integrate_* compute; quad.jl:49,117,155,309), no FMA
       contraction: one thread per integral for large batches; for a few 2D/3D integrals (and
       long 1D grids) one CTA per integral evaluates the terms in parallel and adds them in the
       same order — bit-identical results either way.
       Double64 (FTS_DD64) is the exception: nothing pins its summation order (the reference holds
       no Double64 value; tolerance 1e-28 against a 1e-32 format), so 1D batches are summed by
       groups of 4 lanes per integral (lane g adds the nodes g, g+4, ... of a level in index order,
       the four partial sums are added in lane order; FTS_B200_DD_GROUP=1 restores one thread per
       integral) and levels past 7 by a per-tile tree.  The result depends on neither the batch
       nor the launch.  Double64 2D/3D integrals run one thread each in the reference's order. */
    FTS_ORDER_REFERENCE = 0,
    /* reassociated + FMA, compensated merges (the analogue of the `_avx`/@turbo twins,
       integrate1D.jl:124-147, adaptive.jl:94-133 ...): cooperative CTA/grid per integral. */
    FTS_ORDER_FAST = 1,
    /* library picks per launch from batch size and per-integral work. */
    FTS_ORDER_AUTO = 2
} fts_order;

/* ---- option bits for launches ------------------------------------------------------------ */
enum {
    /* quad()/quad_cmpl() edge rules (quad.jl:41-46, 96-115, 129-153, 303-308):
       any low==up -> exactly 0; flipped bounds -> swap and negate. Without this bit the typed
       adaptive_integrate_* semantics apply (bounds used as given). */
    FTS_OPT_QUAD_EDGE_RULES = 1,
    /* Forced depth (additive; not in the reference): the stop rule is only applied at level ==
       max_levels, so every level is computed.  The reference's forced-depth idiom rtol = atol = 0
       (benchmark/benchmarks.jl:95-110) still stops as soon as two levels agree to the last bit. */
    FTS_OPT_FORCE_DEPTH = 2
};

/* ---- per-integral result flags ----------------------------------------------------------- */
enum {
    FTS_FLAG_CONVERGED = 1,   /* stop rule met (adaptive.jl:66-68) */
    FTS_FLAG_MAX_LEVELS = 2,  /* reached max_levels -> shim emits the reference's @warn (adaptive.jl:71-73) */
    FTS_FLAG_ZERO_WIDTH = 4,  /* quad edge rule returned 0 */
    FTS_FLAG_FLIPPED = 8      /* quad edge rule negated the result */
};

/* ---- integrand signature classes ---------------------------------------------------------- */
typedef enum {
    FTS_KIND_1D = 1,      /* f(x)                 integrate1D*, adaptive_integrate_1D*, quad          */
    FTS_KIND_2D = 2,      /* f(x,y)               integrate2D*, adaptive_integrate_2D*                */
    FTS_KIND_3D = 3,      /* f(x,y,z)             integrate3D*, adaptive_integrate_3D*                */
    FTS_KIND_1D_CMPL = 4, /* f(x, b-x, x-a)       adaptive_integrate_1D_cmpl*, quad_cmpl              */
    FTS_KIND_2D_CMPL = 5, /* f(x,y, b-x,x-a, d-y,y-c)  tensor extension used by BASELINE config C    */
    /* f(const T* x) on a D-dimensional box, 1 <= D <= FTS_ND_MAX: kind = FTS_KIND_ND + D.  The generic n-D tensor
       product the reference's paper names as its next step (JOSS_paper/paper.md:121) — an extension: NVRTC
       integrands only, fixed grids (fts_integrate_batch), f32/f64, reassociated (FAST) summation. */
    FTS_KIND_ND = 16
} fts_kind;

/* ---- builtin integrand families ----------------------------------------------------------- */
/* One statically compiled sm_100a kernel set per family.  `p` = the per-integral parameter row.
   Sources of the formulas: benchmark/benchmarks.jl:32-46, test/families_{1d,2d,3d}.jl,
   test/runtests.jl:58-94, README.md:96-106. */
typedef enum {
    /* 1D: f(x) */
    FTS_F1_POLY7 = 0,        /* sum_{k=0..7} p[k] x^k (Horner)            nparams 8 */
    FTS_F1_EXP = 1,          /* exp(x)                                    nparams 0 */
    FTS_F1_GAUSS = 2,        /* exp(-p0*x^2)                              nparams 1 */
    FTS_F1_RUNGE = 3,        /* 1/(1+p0*x^2)                              nparams 1 */
    FTS_F1_LOG1MX = 4,       /* log(1-x)                                  nparams 0 */
    FTS_F1_LOG1PX = 5,       /* log(1+x)                                  nparams 0 */
    FTS_F1_INVSQRT1MX2 = 6,  /* 1/sqrt(1-x^2)                             nparams 0 */
    FTS_F1_SIN2 = 7,         /* sin(p0*x)^2                               nparams 1 */
    FTS_F1_POLY6 = 8,        /* x^6 - 2x^3 + 0.5                          nparams 0 */
    FTS_F1_INVSQRTABS = 9,   /* p0/sqrt(|x|)                              nparams 1 */
    /* 1D complement: f(x, bmx, xma) */
    FTS_C1_INVSQRT = 20,     /* p0/sqrt(bmx*xma)                          nparams 1 */
    FTS_C1_LOG_BMX = 21,     /* log(bmx)                                  nparams 0 */
    FTS_C1_LOG_XMA = 22,     /* log(xma)                                  nparams 0 */
    FTS_C1_EXP_INV_BMX = 23, /* exp(-1/bmx)                               nparams 0 */
    FTS_C1_EXP_PLUS = 24,    /* exp(x)+bmx+xma  (test/runtests.jl:246)    nparams 0 */
    /* 2D: f(x,y) */
    FTS_F2_POLY2 = 40,       /* p0+p1x+p2y+p3x^2+p4xy+p5y^2               nparams 6 */
    FTS_F2_EXP = 41,         /* exp(x+y)                                  nparams 0 */
    FTS_F2_X2PY2 = 42,       /* x^2+y^2                                   nparams 0 */
    FTS_F2_INVSQRT = 43,     /* 1/sqrt((1-x^2)(1-y^2))                    nparams 0 */
    FTS_F2_INVSQRTABS = 44,  /* 1/sqrt(|x*y|)                             nparams 0 */
    /* 2D complement: f(x,y,bmx,xma,dmy,ymc) */
    FTS_C2_INVSQRT_BMX_YMC = 50, /* 1/sqrt(bmx*ymc) = 1/sqrt((b-x)(y-c))  nparams 0 */
    FTS_C2_INVSQRT_ALL = 51,     /* 1/sqrt(bmx*xma*dmy*ymc)               nparams 0 */
    /* 3D: f(x,y,z) */
    FTS_F3_POLY2 = 60,       /* p0+p1x^2+p2y+p3z^2+p4xyz+p5x^2y^2z^2      nparams 6 */
    FTS_F3_EXP = 61,         /* exp(x+y+z)                                nparams 0 */
    FTS_F3_X2Y2Z2 = 62,      /* x^2*y^2*z^2                               nparams 0 */
    FTS_F3_INVSQRT = 63,     /* 1/sqrt((1-x^2)(1-y^2)(1-z^2))             nparams 0 */
    FTS_F3_RAT = 64,         /* 1/((1+25x^2)(1+16y^2)(1+9z^2))            nparams 0 */
    FTS_F3_LOG = 65,         /* log(1-x)log(1-y)log(1-z)                  nparams 0 */
    FTS_F3_INVSQRTABS = 66   /* 1/sqrt(|x*y*z|)                           nparams 0 */
} fts_family;

/* kind and nparams of a builtin family (host-only helper, no device needed). */
int fts_family_info(int family, int* kind, int* nparams);

/* ---- opaque handles ------------------------------------------------------------------------ */
typedef struct fts_tables_s* fts_tables;       /* fixed grid (x,w,h)            core.jl:256-267   */
typedef struct fts_cache_s* fts_cache;         /* adaptive level tables         caches.jl:31-39, 125-137 */
typedef struct fts_integrand_s* fts_integrand; /* builtin family or NVRTC module */

/* ---- lifecycle ----------------------------------------------------------------------------- */
int fts_abi_version(void);
/* Bind the calling process to CUDA device `device` (one process per GPU). Fails with
   FTS_ERR_NO_DEVICE when no CUDA device is present: there is no CPU fallback. */
int fts_init(int device);
int fts_shutdown(void);
int fts_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor, size_t* global_mem);
/* Copies the calling thread's last error message; returns the full message length. */
size_t fts_last_error(char* buf, size_t len);
/* Pinned, device-mapped host memory (cudaMallocHost).  Optional — every host entry point accepts
   any host pointer — but per-integral operands and result arrays that live in pinned memory
   (from here or cudaHostRegister) are read and written IN PLACE by the kernels over PCIe, with no
   staging copy; pageable buffers are staged through device scratch (chunked copy pipeline). */
int fts_host_alloc(void** ptr, size_t bytes);
int fts_host_free(void* ptr);

/* ---- host-side node generation (cold path; mirrors the reference formulas with libm) -------- */
/* tmax / t_x_max / t_w_max  (core.jl:197-230). which: 0 = tmax(T,D), 1 = t_x_max(T), 2 = t_w_max(T,D) */
int fts_window(int dtype, int which, int D, void* tm_out);
/* tanhsinh(T, N; D)  (core.jl:256-267): n = N div 2 nodes; x_out,w_out hold n elements, h_out one. */
int fts_tanhsinh(int dtype, int N, int D, void* x_out, void* w_out, void* h_out);
/* hopt(T, N, d) = (2/N) W(2dN), W = Lambert W (core.jl:289-291); hmax(T, N, D) = tmax(T,D)/(N div 2)
   (core.jl:293-302).  One element of the dtype is written. */
int fts_hopt(int dtype, int N, double d, void* h_out);
int fts_hmax(int dtype, int N, int D, void* h_out);
/* tanhsinh_opt(T, N, d; D) (core.jl:276-287): nodes at t = h, 2h, ... <= tmax(T,D) with h = hopt.
   *count receives the number of nodes; at most `capacity` of them are written (call with
   capacity 0 to size the buffers). */
int fts_tanhsinh_opt(int dtype, int N, double d, int D, int64_t capacity,
                     void* x_out, void* w_out, void* h_out, int64_t* count);
/* _build_adaptive_1d_cache (caches.jl:59-91): flat level arrays of 2^(L+1)-2 elements
   (level l at offset 2^l-2, l=1..L), complement!=0 selects the t_w_max window. */
int fts_cache1d_generate(int dtype, int max_levels, int complement, void* tm_out,
                         void* initial_x, void* initial_w, void* initial_c,
                         void* xs_flat, void* ws_flat, void* cs_flat);
/* _build_adaptive_tensor_cache (caches.jl:148-188): flat level arrays of 2^(L+2)-4 elements
   (level l holds ALL 2^(l+1) nodes at offset 2^(l+1)-4). */
int fts_cache_tensor_generate(int dtype, int D, int max_levels, void* tm_out,
                              void* initial_x, void* initial_w, void* xs_flat, void* ws_flat);
/* 2D tensor-complement tables (extension for BASELINE config C; NOT in the reference, whose
   quad_cmpl is 1D only, quad.jl:298-318): nodes t=i*h_l as caches.jl:173-177 on the window
   t_w_max(T,2) (core.jl:207-221 with D=2, so products of two weights/complements stay above
   floatmin), plus complements as caches.jl:84.  f32/f64. */
int fts_cache_tensor_cmpl_generate(int dtype, int max_levels, void* tm_out, void* initial_x,
                                   void* initial_w, void* initial_c, void* xs_flat, void* ws_flat,
                                   void* cs_flat);

/* ---- table upload (exact node parity: the caller's own nodes are used verbatim) ------------ */
/* (x,w,h) = tanhsinh(T,N)  — consumed by integrate{1,2,3}D[_avx] (integrate1D.jl:74-147 ...) */
int fts_tables_fixed_upload(int dtype, int n, const void* x, const void* w, const void* h,
                            fts_tables* out);
int fts_tables_free(fts_tables t);
/* _Adaptive1DCache{T} (caches.jl:31-39). kind: 0 regular, 1 complement. */
int fts_cache1d_upload(int dtype, int kind, int max_levels, const void* tm,
                       const void* initial_x, const void* initial_w, const void* initial_c,
                       const void* xs_flat, const void* ws_flat, const void* cs_flat,
                       fts_cache* out);
/* _AdaptiveTensorCache{T} (caches.jl:125-137); scratch vectors xp..zm are NOT part of the
   handle (they live in shared memory per CTA, removing the reference's shared-scratch race). */
int fts_cache_tensor_upload(int dtype, int D, int max_levels, const void* tm,
                            const void* initial_x, const void* initial_w,
                            const void* xs_flat, const void* ws_flat, fts_cache* out);
int fts_cache_tensor_cmpl_upload(int dtype, int max_levels, const void* tm, const void* initial_x,
                                 const void* initial_w, const void* initial_c, const void* xs_flat,
                                 const void* ws_flat, const void* cs_flat, fts_cache* out);
/* length(cache.xs) (caches.jl:25-29 depth guard) and D (0 for 1D caches) */
/* ---- on-device generation ------------------------------------------------------
 * The same tables as fts_tanhsinh / fts_cache*_generate (core.jl:126-185, 256-267;
 * caches.jl:59-91, 148-188), computed by one thread per node straight into device memory: no host
 * arrays, no upload.  Float32/Float64 use CUDA's libm, Double64 double-double series; values agree
 * with the host generators to a few ulp (not bit for bit: different libm).
 * fts_cache_create: D = 0 -> 1D cache (kind 0 regular, 1 complement); D = 2, 3 -> tensor cache;
 * D = 2 with kind 1 -> the 2D complement table. */
int fts_tables_fixed_create(int dtype, int N, int D, fts_tables* out);
int fts_cache_create(int dtype, int D, int kind, int max_levels, fts_cache* out);
/* Read a table back in the layout the *_upload calls take (any output pointer may be null). */
int fts_tables_download(fts_tables t, int* n, void* x, void* w, void* h);
int fts_cache_download(fts_cache c, void* tm, void* initial_x, void* initial_w, void* initial_c,
                       void* xs_flat, void* ws_flat, void* cs_flat);
int fts_cache_info(fts_cache c, int* dtype, int* D, int* kind, int* max_levels);
int fts_cache_free(fts_cache c);

/* ---- integrands ----------------------------------------------------------------------------- */
int fts_integrand_builtin(int family, fts_integrand* out);
/* NVRTC: `body` is the CUDA-C++ body of
 *    template<class T> __device__ T f(ARGS, const T* p)
 * with ARGS per `kind`: (T x) | (T x,T y) | (T x,T y,T z) | (T x,T bmx,T xma) |
 * (T x,T y,T bmx,T xma,T dmy,T ymc).  Compiled for sm_100a with --fmad=false; cached by hash.
 * On failure returns FTS_ERR_NVRTC and copies the compiler log into `log`. */
int fts_integrand_nvrtc(const char* body, int kind, int nparams, fts_integrand* out,
                        char* log, size_t loglen);
int fts_integrand_info(fts_integrand f, int* kind, int* nparams, int* is_builtin);
/* cubin size / kernel count of an NVRTC integrand (compilation needs no GPU) */
int fts_integrand_cubin_size(fts_integrand f, size_t* bytes, int* nkernels);
int fts_integrand_free(fts_integrand f);

/* Pointwise values out[b] = f(x[b], params[b]) of an integrand of kind FTS_KIND_1D, evaluated on the
 * device with the arithmetic of the FAST kernels (HOST buffers, synchronous; f32/f64).  Not in the
 * reference (there f is a Julia callable one can simply call): it lets a caller check an NVRTC body
 * or a builtin family — the tests hold the table-driven exp/log of the builtin families to their ulp
 * bounds with it. */
int fts_eval_batch(fts_integrand f, int dtype, const void* x, const void* params, int params_stride,
                   int64_t n, void* out_value);

/* ---- launches: fixed grids ----------------------------------------------------------------- */
/* integrate{1,2,3}D(f, low, up, x, w, h) / integrate{1,2,3}D_avx
 * (integrate1D.jl:90-102,138-147; integrate2D.jl:91-124,46-73; integrate3D.jl:150-207,68-123).
 * D is taken from the integrand kind. order: FTS_ORDER_REFERENCE = scalar twins,
 * FTS_ORDER_FAST = _avx twins. */
int fts_integrate_batch(fts_tables t, fts_integrand f, int order,
                        const void* low, const void* up, int bounds_stride,
                        const void* params, int params_stride,
                        int64_t batch, void* out_value);
int fts_integrate_batch_dev(fts_tables t, fts_integrand f, int order,
                            const void* low, const void* up, int bounds_stride,
                            const void* params, int params_stride,
                            int64_t batch, void* out_value, void* stream);

/* integrate1D_cmpl(T, f, N) (integrate1D.jl:49-65; unexported in the reference, tested at
 * test/families_1d.jl:20): integral over [-1, 1] of f(x, 1-x, 1+x) on the fixed grid `t`, the
 * complements formed by subtraction as the reference does.  One integral per parameter row
 * (`batch` rows; a family without parameters gives `batch` copies).  f must be of kind
 * FTS_KIND_1D_CMPL.  f32/f64. */
int fts_integrate1d_cmpl_batch(fts_tables t, fts_integrand f, int order, const void* params,
                               int params_stride, int64_t batch, void* out_value);
int fts_integrate1d_cmpl_batch_dev(fts_tables t, fts_integrand f, int order, const void* params,
                                   int params_stride, int64_t batch, void* out_value, void* stream);

/* ---- launches: adaptive -------------------------------------------------------------------- */
/* adaptive_integrate_{1D,2D,3D,1D_cmpl}[_avx] (adaptive.jl:28-75, 94-133, 154-224, 242-333,
 * 352-471, 489-706, 727-785, 805-855), dimension/complement taken from the integrand kind,
 * which must match the cache (1D regular / 1D complement / tensor D).
 * rtol, atol: resolved tolerances (core.jl:32-44 is done by the shim), pointers to one dtype
 * element.  Fails with FTS_ERR_INVALID_ARGUMENT if max_levels exceeds the cache depth
 * (caches.jl:25-29) or a tolerance is negative (core.jl:41-42).
 * Returns old_res (deepest level) with FTS_FLAG_MAX_LEVELS when not converged (adaptive.jl:69-74). */
int fts_adaptive_batch(fts_cache c, fts_integrand f, int order, int options,
                       const void* low, const void* up, int bounds_stride,
                       const void* params, int params_stride,
                       const void* rtol, const void* atol, int max_levels, int64_t batch,
                       void* out_value, void* out_err, int32_t* out_levels, int32_t* out_flags);
int fts_adaptive_batch_dev(fts_cache c, fts_integrand f, int order, int options,
                           const void* low, const void* up, int bounds_stride,
                           const void* params, int params_stride,
                           const void* rtol, const void* atol, int max_levels, int64_t batch,
                           void* out_value, void* out_err, int32_t* out_levels, int32_t* out_flags,
                           void* stream);

/* ---- launches: one large 2D/3D integral sharded over devices ------------------------------ */
/* Partial sum of refinement level `level` (0 = base grid, adaptive.jl:406-418; >=1 = the new
 * points of adaptive.jl:430-456 / 585-690) of ONE adaptive 2D/3D integral for shard `rank` of
 * `nranks`: the flat index over the level's new cells is dealt cyclically over ranks x CTAs
 * (equal work per rank whatever the odd/even class mix).  low/up/params are DEVICE pointers to one
 * box / one parameter row.  Writes the compensated partial (s, c) — two dtype elements — to the
 * DEVICE buffer `partial_dev` on `stream`.  If `state_dev` is non-null and says "converged" the
 * launch exits immediately, so all levels can be queued without a host round trip.
 * The caller exchanges the partials of all ranks (NCCL all-gather of 2 elements) and calls
 * fts_adaptive_sharded_step_dev. */
int fts_adaptive_sharded_level_dev(fts_cache c, fts_integrand f,
                                   const void* low, const void* up, const void* params,
                                   int level, int rank, int nranks,
                                   void* partial_dev, const void* state_dev, void* stream);
/* Stop test after the exchange (adaptive.jl:458-465 / 693-700), one device thread, identical
 * arithmetic on every rank.  partials_dev = [nranks][2] in rank order; state_dev = 8 dtype
 * elements {s_total.s, s_total.c, old_res, value, err_est, done, level_used, h}
 * (done: 0 running, 1 converged). */
int fts_adaptive_sharded_step_dev(fts_cache c, const void* low, const void* up,
                                  const void* rtol, const void* atol, int level, int nranks,
                                  const void* partials_dev, void* state_dev, void* stream);
/* the same with launch option bits; max_levels tells FTS_OPT_FORCE_DEPTH which level is the last */
int fts_adaptive_sharded_step_opts_dev(fts_cache c, const void* low, const void* up,
                                       const void* rtol, const void* atol, int level, int nranks,
                                       const void* partials_dev, void* state_dev, int options,
                                       int max_levels, void* stream);

/* Fused variant ("one kernel does the compute and the exchange"): ONE call queues every level of
 * the integral.  The epilogue of each level kernel stores this rank's (s, c) and then an epoch
 * flag straight into every rank's MAILBOX over NVLink peer memory (cudaIpc-mapped device
 * buffers), and the device-side stop test of every rank waits for the flags in its own mailbox
 * (timeout FTS_B200_PEER_TIMEOUT_MS, default 10 s -> state done = 3).  No NCCL call and no host involvement between levels; levels
 * below `shard_from_level` are too small to shard and are computed whole on every rank.
 * mailboxes[r] = rank r's mailbox mapped into this process (fts_peer_alloc on the owner,
 * fts_peer_open elsewhere; size fts_peer_mailbox_bytes); epoch = 1, 2, 3 ... incremented by the
 * caller once per call, identical on all ranks. */
int fts_peer_mailbox_bytes(int nranks, size_t* bytes);
int fts_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out /* 64 bytes */);
int fts_peer_open(const void* ipc_handle /* 64 bytes */, void** dev_ptr);
int fts_peer_close(void* dev_ptr);
int fts_peer_free(void* dev_ptr);
int fts_adaptive_sharded_fused_dev(fts_cache c, fts_integrand f, const void* low, const void* up,
                                   const void* params, const void* rtol, const void* atol,
                                   int max_levels, int shard_from_level, int rank, int nranks,
                                   void* const* mailboxes, uint64_t epoch, void* state_dev,
                                   void* stream);
/* the same with launch option bits (FTS_OPT_QUAD_EDGE_RULES, FTS_OPT_FORCE_DEPTH) */
int fts_adaptive_sharded_fused_opts_dev(fts_cache c, fts_integrand f, const void* low, const void* up,
                                        const void* params, const void* rtol, const void* atol,
                                        int max_levels, int shard_from_level, int rank, int nranks,
                                        void* const* mailboxes, uint64_t epoch, void* state_dev,
                                        int options, void* stream);

/* Device-side barrier of the ranks' streams through the mailboxes (kernel on `stream`, no host
 * synchronisation): returns at once; work queued behind it starts within microseconds of the same
 * moment on every rank.  `count` = 1, 2, 3 ... per barrier of the group, the same on every rank. */
int fts_peer_barrier_dev(void* const* mailboxes, int rank, int nranks, uint64_t count, void* stream);

/* Per-level exchange wait of the LAST fts_adaptive_sharded_fused_dev call of this process: out[l] =
 * nanoseconds (%globaltimer) the device-side stop test of level l spent spinning on the slowest
 * peer's flag (0 for levels computed whole on every rank), for l < 32; out[32 + l] = %globaltimer
 * at the end of level l's stop test (neighbour differences = per-level durations, exchange
 * included).  Synchronises the device.  The spin gives up after FTS_B200_PEER_TIMEOUT_MS
 * (environment, default 10000).  By default the wait and the stop test run in the last CTA of the
 * level kernel, and the levels that fit one CTA are walked in ONE launch; FTS_B200_STEP_FUSED=0
 * selects the two-launch form (level kernel + one-warp stop-test kernel). */
int fts_peer_wait_ns(int nlevels, uint64_t* out);

/* ---- measurement helpers -------------------------------------------------------------------- */
/* Register-resident DFMA (dtype FTS_F64) / FFMA (FTS_F32) chain microbenchmark: the measured
 * pipe peak used as the roofline denominator (MEASURED_PEAKS.json has no FP64 figure). */
int fts_measure_fma_peak(int dtype, int iters, double* tflops_out, double* ms_out);
/* number of kernels this library launched since the last reset (bench `gpu_launches`) */
int64_t fts_launch_count(int reset);
/* total integrand evaluations of the last adaptive/fixed launch is derivable from levels;
   helper: evaluations of one integral that stopped at `level`. */
int64_t fts_evals_for_level(int D, int level);
int64_t fts_evals_fixed(int D, int n);

#ifdef __cplusplus
}
#endif
#endif /* FTS_B200_H */
