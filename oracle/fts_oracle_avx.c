/*
 * fts_oracle_avx.c — TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * CPU twin of the reference's LoopVectorization (`@turbo`) kernels for the two headline 1D
 * workloads, used ONLY as the timed CPU baseline of bench.py (`cpu_baseline.value_avx`,
 * `--impl reference`): BASELINE.json's north_star asks for "the reference's LoopVectorization CPU
 * path timed on the box's own host cores".  Julia is not in this image, so the `@turbo` loops are
 * restated in C and compiled the way LoopVectorization compiles them:
 *
 *   - the level loop is a SIMD reduction (`#pragma omp simd reduction(+:s)`: lane-wise partial
 *     sums, i.e. reassociated — adaptive.jl:115-118, integrate1D.jl:141-145);
 *   - `exp` is the vector math library's (glibc libmvec `_ZGVeN8v_exp` / `_ZGVdN4v_exp` instead
 *     of SLEEFPirates; both are <= 1-2 ulp vector kernels);
 *   - FMA contraction is allowed (`@turbo` emits `vfmadd`).
 *
 * Compiled by oracle/Makefile with -O3 -march=native -ffast-math -fopenmp (gcc only vectorises
 * libm calls under -ffast-math).  Results are checked against the scalar oracle to the
 * reference's own avx-vs-scalar tolerance (tests/test_oracle_golden.py); nothing in the product
 * loads this file.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/fts_b200.h" /* family ids only */

int fts_or_avx_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* vector width the compiler was allowed to use (reported in bench.py's cpu_baseline.sample) */
int fts_or_avx_vector_bits(void)
{
#if defined(__AVX512F__)
    return 512;
#elif defined(__AVX2__)
    return 256;
#else
    return 128;
#endif
}

#define HALF_PI 1.5707963267948966

/* integrand bodies the compiler inlines into the SIMD loops (Julia specialises on f::F) */
static inline double f_gauss(double x, double a) { return exp(-a * (x * x)); }   /* README.md:102 */
static inline double f_exp(double x, double a) { (void)a; return exp(x); }       /* README.md:174-181 */

/* adaptive_integrate_1D_avx typed core   /root/reference/src/adaptive.jl:94-133 */
#define ADAPTIVE1D_AVX(NAME, FN)                                                                    \
    static inline double NAME(double alpha, double a, double b, double rtol, double atol,           \
                              int max_levels, double tm, const double* ix, const double* iw,        \
                              const double* xs, const double* ws, int* level_out, long* evals_out)  \
    {                                                                                               \
        const double dx = 0.5 * (b - a), x0 = 0.5 * (b + a);                                        \
        double h = tm * 0.5;                                                                        \
        double s_total = HALF_PI * FN(x0, alpha);                                                   \
        for (int i = 0; i < 2; ++i) s_total += iw[i] * (FN(dx * ix[i] + x0, alpha) + FN(x0 - dx * ix[i], alpha)); \
        double old_res = dx * h * s_total, err = 0;                                                 \
        long evals = 5;                                                                             \
        int level = 0;                                                                              \
        for (level = 1; level <= max_levels; ++level) {                                             \
            h *= 0.5;                                                                               \
            const long n = 1L << level;                                                             \
            const double* restrict xl = xs + (n - 2);                                               \
            const double* restrict wl = ws + (n - 2);                                               \
            double s_new = 0;                                                                       \
            _Pragma("omp simd reduction(+ : s_new)")                                                \
            for (long i = 0; i < n; ++i) {                                                          \
                const double xk = xl[i];                                                            \
                s_new += wl[i] * (FN(dx * xk + x0, alpha) + FN(x0 - dx * xk, alpha));               \
            }                                                                                       \
            evals += 2 * n;                                                                         \
            s_total += s_new;                                                                       \
            const double new_res = dx * h * s_total;                                                \
            err = fabs(new_res - old_res);                                                          \
            old_res = new_res;                                                                      \
            const double r = rtol * fabs(new_res);                                                  \
            if (err <= (atol > r ? atol : r)) break;                                                \
        }                                                                                           \
        *level_out = level > max_levels ? max_levels : level;                                       \
        *evals_out = evals;                                                                         \
        return old_res;                                                                             \
    }

ADAPTIVE1D_AVX(adaptive1d_avx_gauss, f_gauss)
ADAPTIVE1D_AVX(adaptive1d_avx_exp, f_exp)

/* The reference's usage loop (README.md:96-106: one adaptive call per parameter over a shared cache)
 * with the `_avx` kernel, OpenMP over the batch.  fam: FTS_F1_GAUSS (params = alpha) or FTS_F1_EXP. */
int fts_or_avx_batch_adaptive1d_f64(int fam, const double* params, int pstride, const double* low,
                                    const double* up, int bstride, double rtol, double atol,
                                    int max_levels, double tm, const double* ix, const double* iw,
                                    const double* xs, const double* ws, long batch, double* value,
                                    int* level, long* evals, int nthreads)
{
    if (fam != FTS_F1_GAUSS && fam != FTS_F1_EXP) return 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1024)
#endif
    for (long b = 0; b < batch; ++b) {
        int lv;
        long ev;
        const double al = params ? params[b * pstride] : 0.0;
        const double v = fam == FTS_F1_GAUSS
                             ? adaptive1d_avx_gauss(al, low[b * bstride], up[b * bstride], rtol, atol, max_levels, tm, ix, iw, xs, ws, &lv, &ev)
                             : adaptive1d_avx_exp(al, low[b * bstride], up[b * bstride], rtol, atol, max_levels, tm, ix, iw, xs, ws, &lv, &ev);
        value[b] = v;
        if (level) level[b] = lv;
        if (evals) evals[b] = ev;
    }
    return 0;
}

/* integrate1D_avx(f, low, up, x, w, h)   /root/reference/src/integrate1D.jl:138-147 */
int fts_or_avx_batch_integrate1d_f64(int fam, const double* params, int pstride, const double* low,
                                     const double* up, int bstride, const double* restrict x,
                                     const double* restrict w, int n, double h, long batch,
                                     double* value, int nthreads)
{
    if (fam != FTS_F1_GAUSS && fam != FTS_F1_EXP) return 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (long b = 0; b < batch; ++b) {
        const double lo = low[b * bstride], hi = up[b * bstride];
        const double dx = 0.5 * (hi - lo), x0 = 0.5 * (hi + lo);
        const double al = params ? params[b * pstride] : 0.0;
        double s = 0;
        if (fam == FTS_F1_GAUSS) {
#pragma omp simd reduction(+ : s)
            for (int i = 0; i < n; ++i) s += w[i] * (f_gauss(x0 - dx * x[i], al) + f_gauss(dx * x[i] + x0, al));
            s += HALF_PI * f_gauss(x0, al);
        } else {
#pragma omp simd reduction(+ : s)
            for (int i = 0; i < n; ++i) s += w[i] * (f_exp(x0 - dx * x[i], al) + f_exp(dx * x[i] + x0, al));
            s += HALF_PI * f_exp(x0, al);
        }
        value[b] = dx * h * s;
    }
    return 0;
}
