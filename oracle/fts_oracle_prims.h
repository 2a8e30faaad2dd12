/*
 * fts_oracle_prims.h — TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed from
 * the product library; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it).
 *
 * Type-generic body of the CPU oracle: a plain-C restatement of the *scalar* (sequential-order)
 * kernels of FastTanhSinhQuadrature.jl v2.1.0.  Included once per element type with
 *     T      element type (float | double | __float128)
 *     ACC    accumulator type (== T: the reference's own rounding; __float128: the exactly
 *            summed value of the same T-rounded terms, used to bound reassociation error)
 *     S(n)   name suffixing macro
 *     M(fn)  libm function of precision T (expf/exp/expq ...)
 *     T_IS_DD  defined when T=__float128 stands in for DoubleFloats.Double64
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Julia does not contract a*b+c into an FMA, so this file must be compiled with
 * -ffp-contract=off (see Makefile).
 */

/* ---------------------------------------------------------------- scalar helpers ---------- */
/* src/core.jl:19-25  _half_pi(T) */
static inline T S(half_pi)(void)
{
#if defined(T_IS_F32)
    return (float)(M_PI / 2); /* Float32(_HALF_PI) */
#elif defined(T_IS_DD)
    return M_PIq * (T)0.5; /* T(pi) * _half(T) */
#else
    return M_PI / 2;
#endif
}

static inline T S(pi)(void)
{
#if defined(T_IS_F32)
    return (float)M_PI;
#elif defined(T_IS_DD)
    return M_PIq;
#else
    return M_PI;
#endif
}

/* prevfloat(one(T)) and floatmin(T) used by the windows (core.jl:197-221). */
static inline T S(prev_one)(void)
{
#if defined(T_IS_F32)
    return nextafterf(1.0f, 0.0f);
#elif defined(T_IS_DD)
    /* Double64 stand-in: 1 - 2^-106 (DoubleFloats.jl is not vendored: parity unpinned). */
    return (T)1 - ldexpq((T)1, -106);
#else
    return nextafter(1.0, 0.0);
#endif
}

static inline T S(floatmin)(void)
{
#if defined(T_IS_F32)
    return FLT_MIN;
#else
    return DBL_MIN; /* Double64 shares Float64's exponent range */
#endif
}

/* src/core.jl:126-139  ordinate(t) = tanh(pi/2 sinh t), clamped to +-prevfloat(1) */
T S(fts_or_ordinate)(T t)
{
    T sinh_t = M(sinh)(t);
    T val = M(tanh)(S(half_pi)() * sinh_t);
    if (val >= (T)1) return S(prev_one)();
    if (val <= -(T)1) return -S(prev_one)();
    return val;
}

/* src/core.jl:172-185  weight(t) = (pi/2 cosh t)/cosh^2(pi/2 sinh t); 0 when |arg|>700 */
T S(fts_or_weight)(T t)
{
    T hp = S(half_pi)();
    T sinh_t = M(sinh)(t), cosh_t = M(cosh)(t);
    T arg = hp * sinh_t;
    if (M(fabs)(arg) > (T)700.0) return (T)0;
    T tmp = M(cosh)(arg);
    T den = tmp * tmp;
#if defined(T_IS_DD)
    if (den > (T)DBL_MAX) return (T)0; /* Double64's cosh^2 overflows like Float64 */
#endif
    return (hp * cosh_t) / den;
}

/* src/core.jl:163-170  ordinate_complement(t) = 2z/(1+z), z = exp(-pi sinh|t|) */
T S(fts_or_ordinate_complement)(T t)
{
    T two = (T)1 + (T)1;
    T u = S(half_pi)() * M(sinh)(M(fabs)(t));
    T z = M(exp)(-two * u);
    return two * z / ((T)1 + z);
}

/* src/core.jl:187-189  inv_ordinate */
T S(fts_or_inv_ordinate)(T t)
{
    return M(asinh)(M(log)(((T)1 + t) / ((T)1 - t)) / S(pi)());
}

/* src/core.jl:197-199 */
T S(fts_or_t_x_max)(void) { return S(fts_or_inv_ordinate)(S(prev_one)()); }

/* src/core.jl:207-221  (calD = T(D), 10 Newton steps) */
T S(fts_or_t_w_max)(int D)
{
    T calD = (T)D;
    T fmin = S(floatmin)();
    T rhs = M(log)(S(pi)()) - M(log)(fmin) / calD;
    T t = M(asinh)(rhs / S(pi)());
    for (int it = 0; it < 10; ++it) {
        T sinh_t = M(sinh)(t), cosh_t = M(cosh)(t);
        T f = S(pi)() * sinh_t - t - rhs;
        T df = S(pi)() * cosh_t - 1;
        t = t - f / df;
    }
    return t;
}

/* src/core.jl:228-230 */
T S(fts_or_tmax)(int D)
{
    T a = S(fts_or_t_x_max)(), b = S(fts_or_t_w_max)(D);
    return a < b ? a : b;
}

/* src/core.jl:256-267  tanhsinh(T,N;D): n = N div 2, h = tm/n, t = range(h, tm, length=n).
 * Julia's range(h, tm, length=n) is a StepRangeLen built with twice-precision arithmetic:
 * t[i] is the correctly rounded value of h + (i-1)*(tm-h)/(n-1) up to that construction, with
 * t[1]==h and t[n]==tm exactly.  We evaluate the same affine map in extended precision and
 * round once (verified against the reference's published values in tests/test_oracle_golden.py). */
void S(fts_or_tanhsinh)(int N, int D, T* x, T* w, T* h_out)
{
    int n = N / 2; /* iseven(N) ? N/2 : (N-1)/2 */
    T tm = S(fts_or_tmax)(D);
    T h = tm / (T)n;
    *h_out = h;
    for (int i = 0; i < n; ++i) {
        T ti;
        if (n == 1) {
            ti = h;
        } else {
            __float128 step = ((__float128)tm - (__float128)h) / (__float128)(n - 1);
            ti = (T)((__float128)h + (__float128)i * step);
            if (i == n - 1) ti = tm;
        }
        /* _ordinate_weight (core.jl:141-156) == (ordinate, weight) of the same t */
        x[i] = S(fts_or_ordinate)(ti);
        w[i] = S(fts_or_weight)(ti);
    }
}

/* src/caches.jl:59-91  _build_adaptive_1d_cache.  Flat layout: level l (1..L) holds 2^l new
 * nodes t=(2i-1)h_l at offset 2^l-2. */
void S(fts_or_cache1d)(int max_levels, int complement, T* tm_out, T* ix, T* iw, T* ic,
                       T* xs, T* ws, T* cs)
{
    T tm = complement ? S(fts_or_t_w_max)(1) : S(fts_or_tmax)(1); /* caches.jl:47-48 */
    T half = (T)1 / ((T)1 + (T)1);
    T h = tm * half;
    *tm_out = tm;
    T h1 = h, h2 = h + h;
    ix[0] = S(fts_or_ordinate)(h1);
    ix[1] = S(fts_or_ordinate)(h2);
    iw[0] = S(fts_or_weight)(h1);
    iw[1] = S(fts_or_weight)(h2);
    ic[0] = S(fts_or_ordinate_complement)(h1);
    ic[1] = S(fts_or_ordinate_complement)(h2);
    for (int level = 1; level <= max_levels; ++level) {
        h *= half;
        long n = 1L << level, off = n - 2;
        for (long i = 1; i <= n; ++i) {
            T tk = (T)(2 * i - 1) * h;
            xs[off + i - 1] = S(fts_or_ordinate)(tk);
            ws[off + i - 1] = S(fts_or_weight)(tk);
            cs[off + i - 1] = S(fts_or_ordinate_complement)(tk);
        }
    }
}

/* src/caches.jl:148-188  _build_adaptive_tensor_cache.  Flat layout: level l holds ALL
 * 2^(l+1) nodes t=i*h_l at offset 2^(l+1)-4. */
void S(fts_or_cache_tensor)(int D, int max_levels, T* tm_out, T* ix, T* iw, T* xs, T* ws)
{
    T tm = S(fts_or_tmax)(D);
    T half = (T)1 / ((T)1 + (T)1);
    T h = tm * half;
    *tm_out = tm;
    T h1 = h, h2 = h + h;
    ix[0] = S(fts_or_ordinate)(h1);
    ix[1] = S(fts_or_ordinate)(h2);
    iw[0] = S(fts_or_weight)(h1);
    iw[1] = S(fts_or_weight)(h2);
    long max_k = 2;
    for (int level = 1; level <= max_levels; ++level) {
        h *= half;
        max_k *= 2;
        long off = max_k - 4;
        for (long i = 1; i <= max_k; ++i) {
            T ti = (T)i * h;
            xs[off + i - 1] = S(fts_or_ordinate)(ti);
            ws[off + i - 1] = S(fts_or_weight)(ti);
        }
    }
}

/* ---------------------------------------------------------------- integrand families ------ */
/* Base.pow_body(x::Float64, n) — compensated power by squaring used by Julia for x^6
 * (Base math.jl; x^2/x^3 literals are plain products). */
static inline T S(jl_pow_int)(T x, int n)
{
#if defined(T_IS_F64)
    double y = 1.0, xnlo = 0.0, ynlo = 0.0, err;
    if (n == 3) return x * x * x;
    while (n > 1) {
        if (n & 1) {
            err = fma(y, xnlo, x * ynlo);
            double yy = x * y;
            ynlo = fma(x, y, -yy);
            y = yy;
            ynlo += err;
        }
        err = x * 2 * xnlo;
        double xx = x * x;
        xnlo = fma(x, x, -xx);
        x = xx;
        xnlo += err;
        n >>= 1;
    }
    err = fma(y, xnlo, x * ynlo);
    return (isfinite(x) && isfinite(err)) ? fma(x, y, err) : x * y;
#elif defined(T_IS_F32)
    double r = 1.0, b = (double)x; /* Float32 path evaluates in Float64 and rounds once */
    for (int k = 0; k < n; ++k) r *= b;
    return (float)r;
#else
    T r = (T)1;
    for (int k = 0; k < n; ++k) r *= x;
    return r;
#endif
}

/* @turbo (LoopVectorization) contracts a*b+c into one FMA inside integrands; the `_avx` golden
 * rows of the reference CSVs were produced that way.  fts_or_set_contract(1) switches the
 * families below to the contracted forms so those rows can be pinned too. */
#ifndef FTS_OR_CONTRACT_DEFINED
#define FTS_OR_CONTRACT_DEFINED
static int g_fts_or_contract = 0;
void fts_or_set_contract(int on) { g_fts_or_contract = on; }
#endif
static inline T S(muladd)(T a, T b, T c) { return g_fts_or_contract ? M(fma)(a, b, c) : a * b + c; }

/* f(x) families; formulas as written in benchmark/benchmarks.jl:33-37, test/families_1d.jl:7-9,
 * README.md:102, test/runtests.jl:74 */
static T S(eval1)(int fam, T x, const T* p)
{
    switch (fam) {
    case FTS_F1_POLY7: {
        T r = p[7];
        for (int k = 6; k >= 0; --k) r = r * x + p[k];
        return r;
    }
    case FTS_F1_EXP: return M(exp)(x);
    case FTS_F1_GAUSS: return M(exp)(-p[0] * (x * x));
    case FTS_F1_RUNGE: return (T)1 / S(muladd)(p[0], x * x, (T)1);
    case FTS_F1_LOG1MX: return M(log)((T)1 - x);
    case FTS_F1_LOG1PX: return M(log)((T)1 + x);
    case FTS_F1_INVSQRT1MX2: return (T)1 / M(sqrt)(S(muladd)(-x, x, (T)1));
    case FTS_F1_SIN2: { T s = M(sin)(p[0] * x); return s * s; }
    case FTS_F1_POLY6: return S(jl_pow_int)(x, 6) - (T)2 * (x * x * x) + (T)0.5;
    case FTS_F1_INVSQRTABS: return p[0] / M(sqrt)(M(fabs)(x));
    default: return (T)NAN;
    }
}

/* f(x, bmx, xma) families; test/families_1d.jl:10,34-35,41, test/runtests.jl:82,246 */
static T S(evalc1)(int fam, T x, T bmx, T xma, const T* p)
{
    switch (fam) {
    case FTS_C1_INVSQRT: return p[0] / M(sqrt)(bmx * xma);
    case FTS_C1_LOG_BMX: return M(log)(bmx);
    case FTS_C1_LOG_XMA: return M(log)(xma);
    case FTS_C1_EXP_INV_BMX: return M(exp)(-(T)1 / bmx);
    case FTS_C1_EXP_PLUS: return M(exp)(x) + bmx + xma;
    default: return (T)NAN;
    }
}

/* f(x,y) families; benchmark/benchmarks.jl:38-41, test/families_2d.jl:5,21,55 */
static T S(eval2)(int fam, T x, T y, const T* p)
{
    switch (fam) {
    case FTS_F2_POLY2:
        return p[0] + p[1] * x + p[2] * y + p[3] * (x * x) + p[4] * (x * y) + p[5] * (y * y);
    case FTS_F2_EXP: return M(exp)(x + y);
    case FTS_F2_X2PY2: return S(muladd)(x, x, y * y);
    case FTS_F2_INVSQRT: return (T)1 / M(sqrt)(S(muladd)(-x, x, (T)1) * S(muladd)(-y, y, (T)1));
    case FTS_F2_INVSQRTABS: return (T)1 / M(sqrt)(M(fabs)(x * y));
    default: return (T)NAN;
    }
}

/* f(x,y,bmx,xma,dmy,ymc): tensor complement extension (BASELINE config C; not in the reference) */
static T S(evalc2)(int fam, T x, T y, T bmx, T xma, T dmy, T ymc, const T* p)
{
    (void)x; (void)y; (void)p;
    switch (fam) {
    case FTS_C2_INVSQRT_BMX_YMC: return (T)1 / M(sqrt)(bmx * ymc);
    case FTS_C2_INVSQRT_ALL: return (T)1 / M(sqrt)(bmx * xma * dmy * ymc);
    default: return (T)NAN;
    }
}

/* f(x,y,z) families; benchmark/benchmarks.jl:42-45, test/families_3d.jl:62-65,76 */
static T S(eval3)(int fam, T x, T y, T z, const T* p)
{
    switch (fam) {
    case FTS_F3_POLY2:
        return p[0] + p[1] * (x * x) + p[2] * y + p[3] * (z * z) + p[4] * (x * y * z) +
               p[5] * ((x * x) * (y * y) * (z * z));
    case FTS_F3_EXP: return M(exp)(x + y + z);
    case FTS_F3_X2Y2Z2: return (x * x) * (y * y) * (z * z);
    case FTS_F3_INVSQRT:
        return (T)1 / M(sqrt)(S(muladd)(-x, x, (T)1) * S(muladd)(-y, y, (T)1) * S(muladd)(-z, z, (T)1));
    case FTS_F3_RAT:
        return (T)1 / (((T)1 + (T)25 * (x * x)) * ((T)1 + (T)16 * (y * y)) * ((T)1 + (T)9 * (z * z)));
    case FTS_F3_LOG: return M(log)((T)1 - x) * M(log)((T)1 - y) * M(log)((T)1 - z);
    case FTS_F3_INVSQRTABS: return (T)1 / M(sqrt)(M(fabs)(x * y * z));
    default: return (T)NAN;
    }
}

/* user callbacks take precedence when non-NULL (ctypes hooks, small cases only) */
typedef T (*S(cb1_t))(T, const T*);
typedef T (*S(cbc1_t))(T, T, T, const T*);
typedef T (*S(cb2_t))(T, T, const T*);
typedef T (*S(cb3_t))(T, T, T, const T*);

typedef struct {
    int fam;
    const T* p;
    void* cb; /* optional callback of the matching arity */
} S(fts_or_fn);

static inline T S(F1)(const S(fts_or_fn)* f, T x)
{
    return f->cb ? ((S(cb1_t))f->cb)(x, f->p) : S(eval1)(f->fam, x, f->p);
}
static inline T S(FC1)(const S(fts_or_fn)* f, T x, T bmx, T xma)
{
    return f->cb ? ((S(cbc1_t))f->cb)(x, bmx, xma, f->p) : S(evalc1)(f->fam, x, bmx, xma, f->p);
}
static inline T S(F2)(const S(fts_or_fn)* f, T x, T y)
{
    return f->cb ? ((S(cb2_t))f->cb)(x, y, f->p) : S(eval2)(f->fam, x, y, f->p);
}
static inline T S(F3)(const S(fts_or_fn)* f, T x, T y, T z)
{
    return f->cb ? ((S(cb3_t))f->cb)(x, y, z, f->p) : S(eval3)(f->fam, x, y, z, f->p);
}

