/*
 * fts_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the fts_b200 parity tests: a plain-C restatement of the scalar kernels of
 * FastTanhSinhQuadrature.jl v2.1.0 (reference at /root/reference, pure Julia; Julia is not
 * installed in this image so the reference itself cannot be executed — see DESIGN.md).
 * Pinned against the reference's published result values (benchmark/results/timings*.csv,
 * column `value`) and the analytic known-answer tests of test/ (*.jl) by tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
 * may load this library.  The product (libfts_b200.so) never links, loads or calls it.
 *
 * Instantiations:
 *   _f32, _f64           element type T, accumulation in T  (reference rounding)
 *   _f32x, _f64x         element type T, accumulation in __float128 (exact sum of same terms)
 *   _dd                  T = __float128 standing in for DoubleFloats.Double64 (parity unpinned:
 *                        the reference's tests hold no Double64 quad_cmpl values)
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <quadmath.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/fts_b200.h" /* family ids only */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* ------------------------------------------------------------------ float ------------------ */
#define T float
#define T_IS_F32
#define S(n) CAT(n, _f32)
#define M(fn) CAT(fn, f)
#include "fts_oracle_prims.h"
#define ACC float
#define ACC_IS_T 1
#define K(n) CAT(n, _f32)
#include "fts_oracle_kernels.h"
#undef ACC
#undef ACC_IS_T
#undef K
#define ACC __float128
#define ACC_IS_T 0
#define K(n) CAT(n, _f32x)
#include "fts_oracle_kernels.h"
#undef ACC
#undef ACC_IS_T
#undef K
#undef T
#undef T_IS_F32
#undef S
#undef M

/* ------------------------------------------------------------------ double ----------------- */
#define T double
#define T_IS_F64
#define S(n) CAT(n, _f64)
#define M(fn) fn
#include "fts_oracle_prims.h"
#define ACC double
#define ACC_IS_T 1
#define K(n) CAT(n, _f64)
#include "fts_oracle_kernels.h"
#undef ACC
#undef ACC_IS_T
#undef K
#define ACC __float128
#define ACC_IS_T 0
#define K(n) CAT(n, _f64x)
#include "fts_oracle_kernels.h"
#undef ACC
#undef ACC_IS_T
#undef K
#undef T
#undef T_IS_F64
#undef S
#undef M

/* ------------------------------------------------------------------ Double64 stand-in ------ */
#define T __float128
#define T_IS_DD
#define S(n) CAT(n, _dd)
#define M(fn) CAT(fn, q)
#include "fts_oracle_prims.h"
#define ACC __float128
#define ACC_IS_T 1
#define K(n) CAT(n, _dd)
#include "fts_oracle_kernels.h"
#undef ACC
#undef ACC_IS_T
#undef K
#undef T
#undef T_IS_DD
#undef S
#undef M

/* ------------------------------------------------------------------ dd <-> (hi,lo) --------- */
void fts_or_dd_split(const __float128* q, double* hi, double* lo, long n)
{
    for (long i = 0; i < n; ++i) {
        hi[i] = (double)q[i];
        lo[i] = (double)(q[i] - (__float128)hi[i]);
    }
}
void fts_or_dd_join(const double* hi, const double* lo, __float128* q, long n)
{
    for (long i = 0; i < n; ++i) q[i] = (__float128)hi[i] + (__float128)lo[i];
}
/* decimal string <-> __float128 for Python */
void fts_or_q_from_str(const char* s, __float128* out) { *out = strtoflt128(s, NULL); }
int fts_or_q_to_str(const __float128* q, char* buf, int len)
{
    return quadmath_snprintf(buf, (size_t)len, "%.36Qe", *q);
}

/* ------------------------------------------------------------------ batch drivers ---------- */
/* Host-core CPU baseline: the reference loops one integral at a time over a shared cache
 * (README.md:96-106); here the same per-integral kernel runs under an OpenMP parallel-for. */
int fts_or_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Julia specialises and inlines the integrand into adaptive_integrate_1D (`f::F ... where F`,
 * adaptive.jl:28); the per-family bodies below give gcc the same chance (constant family id +
 * flatten), so the CPU baseline is not handicapped by the oracle's run-time family switch. */
#define OR_BATCH1D_BODY(FAMCONST)                                                                  \
    for (long b = 0; b < batch; ++b) {                                                             \
        fts_or_result_f64 r;                                                                       \
        fts_or_adaptive1d_f64(FAMCONST, params + b * pstride, NULL, low[b * bstride],              \
                              up[b * bstride], rtol, atol, max_levels, tm, ix, iw, xs, ws, &r);    \
        value[b] = r.value;                                                                        \
        if (err) err[b] = r.err;                                                                   \
        if (level) level[b] = r.level;                                                             \
        if (conv) conv[b] = r.converged;                                                           \
        if (evals) evals[b] = r.evals;                                                             \
    }

__attribute__((flatten)) void fts_or_batch_adaptive1d_f64(
    int fam, const double* params, int pstride, const double* low, const double* up, int bstride,
    double rtol, double atol, int max_levels, double tm, const double* ix, const double* iw,
    const double* xs, const double* ws, long batch, double* value, double* err, int* level,
    int* conv, long* evals, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (fam == FTS_F1_GAUSS) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
        OR_BATCH1D_BODY(FTS_F1_GAUSS)
    } else if (fam == FTS_F1_EXP) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
        OR_BATCH1D_BODY(FTS_F1_EXP)
    } else {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
        OR_BATCH1D_BODY(fam)
    }
}

void fts_or_batch_adaptive1d_f32(int fam, const float* params, int pstride, const float* low,
                                 const float* up, int bstride, float rtol, float atol,
                                 int max_levels, float tm, const float* ix, const float* iw,
                                 const float* xs, const float* ws, long batch, float* value,
                                 float* err, int* level, int* conv, long* evals, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1024)
#endif
    for (long b = 0; b < batch; ++b) {
        fts_or_result_f32 r;
        fts_or_adaptive1d_f32(fam, params + b * pstride, NULL, low[b * bstride], up[b * bstride],
                              rtol, atol, max_levels, tm, ix, iw, xs, ws, &r);
        value[b] = r.value;
        if (err) err[b] = r.err;
        if (level) level[b] = r.level;
        if (conv) conv[b] = r.converged;
        if (evals) evals[b] = r.evals;
    }
}

void fts_or_batch_adaptive1d_cmpl_f64(int fam, const double* params, int pstride, const double* low,
                                      const double* up, int bstride, double rtol, double atol,
                                      int max_levels, double tm, const double* ix, const double* iw,
                                      const double* ic, const double* xs, const double* ws,
                                      const double* cs, long batch, double* value, double* err,
                                      int* level, int* conv, long* evals, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256)
#endif
    for (long b = 0; b < batch; ++b) {
        fts_or_result_f64 r;
        fts_or_adaptive1d_cmpl_f64(fam, params + b * pstride, NULL, low[b * bstride], up[b * bstride],
                                   rtol, atol, max_levels, tm, ix, iw, ic, xs, ws, cs, &r);
        value[b] = r.value;
        if (err) err[b] = r.err;
        if (level) level[b] = r.level;
        if (conv) conv[b] = r.converged;
        if (evals) evals[b] = r.evals;
    }
}

void fts_or_batch_adaptive1d_cmpl_dd(int fam, const __float128* params, int pstride,
                                     const __float128* low, const __float128* up, int bstride,
                                     const __float128* rtol, const __float128* atol, int max_levels,
                                     const __float128* tm, const __float128* ix, const __float128* iw,
                                     const __float128* ic, const __float128* xs, const __float128* ws,
                                     const __float128* cs, long batch, __float128* value,
                                     __float128* err, int* level, int* conv, long* evals, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (long b = 0; b < batch; ++b) {
        fts_or_result_dd r;
        fts_or_adaptive1d_cmpl_dd(fam, params + b * pstride, NULL, low[b * bstride], up[b * bstride],
                                  *rtol, *atol, max_levels, *tm, ix, iw, ic, xs, ws, cs, &r);
        value[b] = r.value;
        if (err) err[b] = r.err;
        if (level) level[b] = r.level;
        if (conv) conv[b] = r.converged;
        if (evals) evals[b] = r.evals;
    }
}

void fts_or_batch_integrate1d_f64(int fam, const double* params, int pstride, const double* low,
                                  const double* up, int bstride, const double* x, const double* w,
                                  int n, double h, long batch, double* value, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (long b = 0; b < batch; ++b)
        value[b] = fts_or_integrate1d_f64(fam, params + b * pstride, NULL, low[b * bstride],
                                          up[b * bstride], x, w, n, h);
}

void fts_or_batch_adaptive2d_f64(int fam, const double* params, int pstride, const double* low,
                                 const double* up, int bstride, double rtol, double atol,
                                 int max_levels, double tm, const double* ix, const double* iw,
                                 const double* xs, const double* ws, long batch, double* value,
                                 double* err, int* level, int* conv, long* evals, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (long b = 0; b < batch; ++b) {
        fts_or_result_f64 r;
        fts_or_adaptive2d_f64(fam, params + b * pstride, NULL, low + b * bstride, up + b * bstride,
                              rtol, atol, max_levels, tm, ix, iw, xs, ws, &r);
        value[b] = r.value;
        if (err) err[b] = r.err;
        if (level) level[b] = r.level;
        if (conv) conv[b] = r.converged;
        if (evals) evals[b] = r.evals;
    }
}

/* Partial sum of one 3D refinement level for shard `rank` of `nranks` (outer index i dealt
 * cyclically; level 0 = the base grid, owned by rank 0): the CPU stand-in used by the gloo
 * world_size-2 tests of the multi-GPU orchestration (tests/test_distributed_cpu.py). */
double fts_or_adaptive3d_level_partial_f64(int fam, const double* p, const double* low, const double* up,
                                           const double* ix, const double* iw, const double* xs,
                                           const double* ws, int level, int rank, int nranks)
{
    fts_or_fn_f64 f = {fam, p, NULL};
    box3_f64 bx;
    bx.dx = 0.5 * (up[0] - low[0]); bx.x0 = 0.5 * (up[0] + low[0]);
    bx.dy = 0.5 * (up[1] - low[1]); bx.y0 = 0.5 * (up[1] + low[1]);
    bx.dz = 0.5 * (up[2] - low[2]); bx.z0 = 0.5 * (up[2] + low[2]);
    bx.w0 = M_PI / 2;
    bx.w02 = bx.w0 * bx.w0;
    if (level == 0) return rank == 0 ? fts_or_adaptive3d_level0_f64(&f, &bx, ix, iw) : 0.0;
    long n = 2L << level, off = n - 4;
    return fts_or_adaptive3d_level_sum_f64(&f, &bx, xs + off, ws + off, n, rank, nranks);
}

/* One 3D level sum with the outer index i split over OpenMP threads (BASELINE.md §3 plan for
 * the single large integral); partials are combined in thread order. */
double fts_or_adaptive3d_level_sum_omp_f64(int fam, const double* p, const double* low,
                                           const double* up, const double* xl, const double* wl,
                                           long n, int nthreads)
{
    fts_or_fn_f64 f = {fam, p, NULL};
    box3_f64 bx;
    bx.dx = 0.5 * (up[0] - low[0]); bx.x0 = 0.5 * (up[0] + low[0]);
    bx.dy = 0.5 * (up[1] - low[1]); bx.y0 = 0.5 * (up[1] + low[1]);
    bx.dz = 0.5 * (up[2] - low[2]); bx.z0 = 0.5 * (up[2] + low[2]);
    bx.w0 = M_PI / 2;
    bx.w02 = bx.w0 * bx.w0;
    int nt = nthreads > 0 ? nthreads : fts_or_num_threads();
    double total = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(static, 1) reduction(+ : total)
#endif
    for (int r = 0; r < nt * 8; ++r) total += fts_or_adaptive3d_level_sum_f64(&f, &bx, xl, wl, n, r, nt * 8);
    return total;
}

/* adaptive_integrate_3D (adaptive.jl:352-471) with every level's new-point sum split over OpenMP
 * threads along the outer index i and accumulated in __float128: the exactly-summed value of the
 * same T-rounded terms as fts_or_adaptive3d_f64x (the order of an exact sum does not matter to
 * ~1e-34), but a level-8 integral (1.08e9 evaluations) takes seconds instead of a minute.  Used by
 * the GPU parity tests at BASELINE config D depths (forced depth L = 7, 8). */
void fts_or_adaptive3d_omp_f64x(int fam, const double* p, const double* low, const double* up,
                                double rtol, double atol, int max_levels, double tm, const double* ix,
                                const double* iw, const double* xs, const double* ws, int nthreads,
                                fts_or_result_f64x* out)
{
    fts_or_fn_f64 f = {fam, p, NULL};
    box3_f64x bx;
    bx.dx = 0.5 * (up[0] - low[0]); bx.x0 = 0.5 * (up[0] + low[0]);
    bx.dy = 0.5 * (up[1] - low[1]); bx.y0 = 0.5 * (up[1] + low[1]);
    bx.dz = 0.5 * (up[2] - low[2]); bx.z0 = 0.5 * (up[2] + low[2]);
    bx.w0 = half_pi_f64();
    bx.w02 = bx.w0 * bx.w0;
    int nt = nthreads > 0 ? nthreads : fts_or_num_threads();
    double h = tm * 0.5;
    __float128 s_total = fts_or_adaptive3d_level0_f64x(&f, &bx, ix, iw);
    long evals = 125;
#define SCALE3X(s) ((double)(((__float128)bx.dx * (__float128)bx.dy * (__float128)bx.dz * ((__float128)h * (__float128)h * (__float128)h)) * (s)))
    double old_res = SCALE3X(s_total);
    double err = 0;
    out->level = 0;
    out->converged = 0;
    for (int level = 1; level <= max_levels; ++level) {
        h *= 0.5;
        long n = 2L << level, off = n - 4;
        int parts = nt * 8;
        if (parts > n) parts = (int)n;
        __float128 partial[parts];
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1)
#endif
        for (int r = 0; r < parts; ++r) partial[r] = fts_or_adaptive3d_level_sum_f64x(&f, &bx, xs + off, ws + off, n, r, parts);
        __float128 s_new = 0;
        for (int r = 0; r < parts; ++r) s_new += partial[r];
        evals += 7 * n * n * n + 9 * n * n + 3 * n;
        s_total += s_new;
        double new_res = SCALE3X(s_total);
        err = fabs(new_res - old_res);
        out->level = level;
        if (err <= err_target_f64x(new_res, rtol, atol)) {
            out->value = new_res;
            out->err = err;
            out->converged = 1;
            out->evals = evals;
            return;
        }
        old_res = new_res;
    }
#undef SCALE3X
    out->value = old_res;
    out->err = err;
    out->evals = evals;
}
