// fts_families.cuh — builtin integrand families (one statically compiled sm_100a kernel set each).
//
// Formulas are written exactly as the reference's benchmark/test integrands
// (/root/reference/benchmark/benchmarks.jl:33-45, test/families_{1d,2d,3d}.jl, README.md:102,
// test/runtests.jl:58-94,246) with Julia's evaluation order; because the library is compiled
// with --fmad=false nothing is contracted unless written as fma().  `FAST` selects the
// contracted forms LoopVectorization's @turbo would emit for the `_avx` twins.
//
// Functor protocol (also what NVRTC user integrands are wrapped into):
//   struct F { static constexpr int NP;  template<class T,bool FAST> static T eval(args..., const T* p); }
#pragma once
#include "fts_math.cuh"

namespace fts {

template <class F> struct is_cmpl2d;

// Base.pow_body(x, 6) (compensated power by squaring) for Float64; plain products otherwise
template <class T> __device__ __forceinline__ T jl_pow6(T x)
{
    T x2 = x * x;
    return (x2 * x2) * x2;
}
template <> __device__ __forceinline__ double jl_pow6<double>(double x)
{
    // n = 6 = 0b110: square, (multiply, square), finish   (Base math.jl pow_body)
    double x2 = x * x, x2lo = ::fma(x, x, -x2);
    double err = x2lo;                 // muladd(1, xnlo, x*0)
    double y = x2, ylo = err;          // two_mul(x2, 1) + err
    err = x2 * 2 * x2lo;
    double x4 = x2 * x2, x4lo = ::fma(x2, x2, -x4);
    x4lo += err;
    err = ::fma(y, x4lo, x4 * ylo);
    double r = ::fma(x4, y, err);
    return (isfinite(x4) && isfinite(err)) ? r : x4 * y;
}
template <> __device__ __forceinline__ float jl_pow6<float>(float x)
{
    double b = (double)x, b2 = b * b;
    return (float)((b2 * b2) * b2);
}

// ---------------------------------------------------------------- 1D: f(x) ------------------
struct F1Poly7 {
    static constexpr int NP = 8;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T* p)
    {
        T r = p[7];
#pragma unroll
        for (int k = 6; k >= 0; --k) r = muladd<T, FAST>(r, x, p[k]);
        return r;
    }
};
struct F1Exp {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    static constexpr bool USES_EXP_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T*) { return fm::exp(x); }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T*, T (&f)[N]) { exp_n<N, MODE>(x, f); }
    template <class T> static __device__ __forceinline__ T exp_arg_bound(T reach, const T*) { return reach; }  // max |x| over the interval
};
struct F1Gauss {  // exp(-α*x^2)   README.md:102
    static constexpr int NP = 1;
    static constexpr bool HAS_EVALN = true;
    static constexpr bool USES_EXP_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T* p) { return fm::exp(-p[0] * (x * x)); }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T* p, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = -p[0] * (x[i] * x[i]);
        exp_n<N, MODE>(a, f);
    }
    template <class T> static __device__ __forceinline__ T exp_arg_bound(T reach, const T* p) { return fm::abs(p[0]) * (reach * reach); }
};
struct F1Runge {  // 1/(1+25x^2)   benchmarks.jl:34
    static constexpr int NP = 1;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T* p)
    {
        return num<T>::one() / muladd<T, FAST>(p[0], x * x, num<T>::one());
    }
};
// log families: fm::log_tab is the table-driven log of fts_math.cuh (Float64; <= 0.72 ulp), evaluated in
// lock-step for the N points a kernel hands over at once
struct F1Log1mx {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    static constexpr bool USES_LOG_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T*) { return fm::log_tab(num<T>::one() - x); }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = num<T>::one() - x[i];
        log_n<N>(a, f);
    }
};
struct F1Log1px {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    static constexpr bool USES_LOG_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T*) { return fm::log_tab(num<T>::one() + x); }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = num<T>::one() + x[i];
        log_n<N>(a, f);
    }
};
struct F1InvSqrt1mx2 {  // 1/sqrt(1-x^2)   benchmarks.jl:36
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T*)
    {
        return fm::inv_sqrt<FAST>(muladd<T, FAST>(-x, x, num<T>::one()));
    }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = muladd<T, FAST>(-x[i], x[i], num<T>::one());
        inv_sqrt_n<FAST>(a, f);
    }
};
struct F1Sin2 {  // sin(1000x)^2   benchmarks.jl:37
    static constexpr int NP = 1;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T* p)
    {
        T s = fm::sin(p[0] * x);
        return s * s;
    }
};
template <> __device__ __forceinline__ dd F1Sin2::eval<dd, false>(dd x, const dd* p) { double s = ::sin((p[0] * x).hi); return dd(s * s); }
template <> __device__ __forceinline__ dd F1Sin2::eval<dd, true>(dd x, const dd* p) { double s = ::sin((p[0] * x).hi); return dd(s * s); }
struct F1Poly6 {  // x^6 - 2x^3 + 0.5   benchmarks.jl:33
    static constexpr int NP = 0;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T*)
    {
        return jl_pow6<T>(x) - num<T>::from_double(2.0) * (x * x * x) + num<T>::half();
    }
};
struct F1InvSqrtAbs {  // c/sqrt(|x|)   test/runtests.jl:74
    static constexpr int NP = 1;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, const T* p)
    {
        if constexpr (sizeof(T) == 16) return p[0] * dd_rsqrt(fm::abs(x));  // Double64: no reference bits to keep, 5x fewer instructions
        else return p[0] / fm::sqrt(fm::abs(x));
    }
};

// ---------------------------------------------------------------- 1D complement: f(x,bmx,xma)
struct C1InvSqrt {  // c/sqrt(bmx*xma)   test/runtests.jl:82, families_1d.jl:10
    static constexpr int NP = 1;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T, T bmx, T xma, const T* p)
    {
        if constexpr (sizeof(T) == 16) return p[0] * dd_rsqrt(bmx * xma);  // Double64: no reference bits to keep, 5x fewer instructions
        else return p[0] / fm::sqrt(bmx * xma);
    }
};
struct C1LogBmx {
    static constexpr int NP = 0;
    static constexpr bool USES_LOG_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T, T bmx, T, const T*) { return fm::log_tab(bmx); }
};
struct C1LogXma {
    static constexpr int NP = 0;
    static constexpr bool USES_LOG_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T, T, T xma, const T*) { return fm::log_tab(xma); }
};
struct C1ExpInvBmx {  // exp(-1/bmx)   families_1d.jl:41
    static constexpr int NP = 0;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T, T bmx, T, const T*) { return fm::exp(-num<T>::one() / bmx); }
};
struct C1ExpPlus {  // exp(x)+bmx+xma   test/runtests.jl:246
    static constexpr int NP = 0;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T bmx, T xma, const T*) { return fm::exp(x) + bmx + xma; }
};

// ---------------------------------------------------------------- 2D: f(x,y) ----------------
struct F2Poly2 {
    static constexpr int NP = 6;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, const T* p)
    {
        return p[0] + p[1] * x + p[2] * y + p[3] * (x * x) + p[4] * (x * y) + p[5] * (y * y);
    }
};
struct F2Exp {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    static constexpr bool USES_EXP_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, const T*) { return fm::exp(x + y); }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T (&y)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = x[i] + y[i];
        exp_n<N, MODE>(a, f);
    }
    template <class T> static __device__ __forceinline__ T exp_arg_bound(T reach, const T*) { return reach; }  // reach = sum over axes of |c|+|d|
};
struct F2X2pY2 {
    static constexpr int NP = 0;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, const T*) { return muladd<T, FAST>(x, x, y * y); }
};
struct F2InvSqrt {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, const T*)
    {
        const T one = num<T>::one();
        return fm::inv_sqrt<FAST>(muladd<T, FAST>(-x, x, one) * muladd<T, FAST>(-y, y, one));
    }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T (&y)[N], const T*, T (&f)[N])
    {
        const T one = num<T>::one();
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = muladd<T, FAST>(-x[i], x[i], one) * muladd<T, FAST>(-y[i], y[i], one);
        inv_sqrt_n<FAST>(a, f);
    }
};
struct F2InvSqrtAbs {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, const T*) { return fm::inv_sqrt<FAST>(fm::abs(x * y)); }
    template <class T, bool FAST, int N, int MODE> static __device__ __forceinline__ void evaln(const T (&x)[N], const T (&y)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = fm::abs(x[i] * y[i]);
        inv_sqrt_n<FAST>(a, f);
    }
};

// ---------------------------------------------------------------- 2D complement -------------
struct C2InvSqrtBmxYmc {  // 1/sqrt((b-x)(y-c))   BASELINE config C
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T, T, T bmx, T, T, T ymc, const T*)
    {
        return fm::inv_sqrt<FAST>(bmx * ymc);
    }
    template <class T, bool FAST, int N, int MODE>
    static __device__ __forceinline__ void evaln(const T (&)[N], const T (&)[N], const T (&bmx)[N], const T (&)[N], const T (&)[N], const T (&ymc)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = bmx[i] * ymc[i];
        inv_sqrt_n<FAST>(a, f);
    }
};
struct C2InvSqrtAll {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T, T, T bmx, T xma, T dmy, T ymc, const T*)
    {
        return fm::inv_sqrt<FAST>(bmx * xma * dmy * ymc);
    }
    template <class T, bool FAST, int N, int MODE>
    static __device__ __forceinline__ void evaln(const T (&)[N], const T (&)[N], const T (&bmx)[N], const T (&xma)[N], const T (&dmy)[N], const T (&ymc)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = bmx[i] * xma[i] * dmy[i] * ymc[i];
        inv_sqrt_n<FAST>(a, f);
    }
};

template <> struct is_cmpl2d<C2InvSqrtBmxYmc> { static constexpr bool value = true; };
template <> struct is_cmpl2d<C2InvSqrtAll> { static constexpr bool value = true; };

// ---------------------------------------------------------------- 3D: f(x,y,z) --------------
struct F3Poly2 {
    static constexpr int NP = 6;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, T z, const T* p)
    {
        return p[0] + p[1] * (x * x) + p[2] * y + p[3] * (z * z) + p[4] * (x * y * z) + p[5] * ((x * x) * (y * y) * (z * z));
    }
};
struct F3Exp {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    static constexpr bool USES_EXP_TAB = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, T z, const T*) { return fm::exp(x + y + z); }
    template <class T, bool FAST, int N, int MODE>
    static __device__ __forceinline__ void evaln(const T (&x)[N], const T (&y)[N], const T (&z)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = x[i] + y[i] + z[i];
        exp_n<N, MODE>(a, f);
    }
    template <class T> static __device__ __forceinline__ T exp_arg_bound(T reach, const T*) { return reach; }
};
struct F3X2Y2Z2 {
    static constexpr int NP = 0;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, T z, const T*) { return (x * x) * (y * y) * (z * z); }
};
struct F3InvSqrt {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, T z, const T*)
    {
        const T one = num<T>::one();
        return fm::inv_sqrt<FAST>(muladd<T, FAST>(-x, x, one) * muladd<T, FAST>(-y, y, one) * muladd<T, FAST>(-z, z, one));
    }
    template <class T, bool FAST, int N, int MODE>
    static __device__ __forceinline__ void evaln(const T (&x)[N], const T (&y)[N], const T (&z)[N], const T*, T (&f)[N])
    {
        const T one = num<T>::one();
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = muladd<T, FAST>(-x[i], x[i], one) * muladd<T, FAST>(-y[i], y[i], one) * muladd<T, FAST>(-z[i], z[i], one);
        inv_sqrt_n<FAST>(a, f);
    }
};
struct F3Rat {
    static constexpr int NP = 0;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, T z, const T*)
    {
        const T one = num<T>::one();
        return one / (muladd<T, FAST>(num<T>::from_double(25.0), x * x, one) * muladd<T, FAST>(num<T>::from_double(16.0), y * y, one) *
                      muladd<T, FAST>(num<T>::from_double(9.0), z * z, one));
    }
};
struct F3Log {  // log(1-x) log(1-y) log(1-z)   benchmarks.jl:45
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    static constexpr bool USES_LOG_TAB = true;
    static constexpr bool HAS_CELLS_RUN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, T z, const T*)
    {
        const T one = num<T>::one();
        return fm::log_tab(one - x) * fm::log_tab(one - y) * fm::log_tab(one - z);
    }
    // N points at once: 3N logs in lock-step (the compiler merges the duplicates — the 8 reflections of
    // a cell have only 2 distinct x, y and z each)
    template <class T, bool FAST, int N, int MODE>
    static __device__ __forceinline__ void evaln(const T (&x)[N], const T (&y)[N], const T (&z)[N], const T*, T (&f)[N])
    {
        const T one = num<T>::one();
        T a[3 * N], l[3 * N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            a[i] = one - x[i];
            a[N + i] = one - y[i];
            a[2 * N + i] = one - z[i];
        }
        log_n<3 * N>(a, l);
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = l[i] * l[N + i] * l[2 * N + i];
    }
    // CH cells along k that share (x+-, y+-): sum_c w[c] * (sum of the 8 reflections of cell c), every value
    // formed exactly as eval() forms it, (log(1-x) log(1-y)) log(1-z), and added in the order of cell3 — but the
    // four x/y logs and their four products are computed once per run instead of once per reflection
    // (what a compiler's common-subexpression elimination does to the reference's k-loop, adaptive.jl:600-607)
    template <class T, int CH>
    static __device__ __forceinline__ T cells_run(T xp, T xm, T yp, T ym, const T (&zp)[CH], const T (&zm)[CH], const T (&w)[CH], const T*)
    {
        const T one = num<T>::one();
        T a[4 + 2 * CH], l[4 + 2 * CH];
        a[0] = one - xp; a[1] = one - xm; a[2] = one - yp; a[3] = one - ym;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            a[4 + c] = one - zp[c];
            a[4 + CH + c] = one - zm[c];
        }
        log_n<4 + 2 * CH>(a, l);
        const T pp = l[0] * l[2], mp = l[1] * l[2], pm = l[0] * l[3], mm = l[1] * l[3];  // (x+,y+) (x-,y+) (x+,y-) (x-,y-)
        T acc = num<T>::zero();
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const T lzp = l[4 + c], lzm = l[4 + CH + c];
            const T f = ((pp * lzp + mp * lzp) + (pm * lzp + mm * lzp)) + ((pp * lzm + mp * lzm) + (pm * lzm + mm * lzm));
            acc = fm::fma(w[c], f, acc);
        }
        return acc;
    }
};
struct F3InvSqrtAbs {
    static constexpr int NP = 0;
    static constexpr bool HAS_EVALN = true;
    template <class T, bool FAST> static __device__ __forceinline__ T eval(T x, T y, T z, const T*) { return fm::inv_sqrt<FAST>(fm::abs(x * y * z)); }
    template <class T, bool FAST, int N, int MODE>
    static __device__ __forceinline__ void evaln(const T (&x)[N], const T (&y)[N], const T (&z)[N], const T*, T (&f)[N])
    {
        T a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = fm::abs(x[i] * y[i] * z[i]);
        inv_sqrt_n<FAST>(a, f);
    }
};

}  // namespace fts
