// fts_tablegen.cpp — host-side tanh-sinh node/weight/complement tables (cold path).
//
// Mirrors the reference's L0/L1 layers (/root/reference/src/core.jl:126-267, src/caches.jl:59-91,
// 148-188) with libm in the element precision; Double64 tables are computed in __float128 and
// split into (hi, lo).  A Julia caller uploads its OWN tables (fts_cache1d_upload ...) so node
// parity with the reference is exact by construction; these generators serve non-Julia hosts.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <quadmath.h>

#include "../../include/fts_b200.h"

namespace {

typedef __float128 q128;

// ---- precision traits -----------------------------------------------------------------------
template <class T> struct P;
template <> struct P<float> {
    static float sinh(float v) { return ::sinhf(v); }
    static float cosh(float v) { return ::coshf(v); }
    static float tanh(float v) { return ::tanhf(v); }
    static float exp(float v) { return ::expf(v); }
    static float log(float v) { return ::logf(v); }
    static float asinh(float v) { return ::asinhf(v); }
    static float abs(float v) { return ::fabsf(v); }
    static float pi() { return (float)M_PI; }
    static float half_pi() { return (float)(M_PI / 2); }           // core.jl:24
    static float below_one() { return ::nextafterf(1.0f, 0.0f); }  // prevfloat(one(T))
    static float tiny() { return FLT_MIN; }                        // floatmin(T)
    static bool sq_overflows(float) { return false; }
};
template <> struct P<double> {
    static double sinh(double v) { return ::sinh(v); }
    static double cosh(double v) { return ::cosh(v); }
    static double tanh(double v) { return ::tanh(v); }
    static double exp(double v) { return ::exp(v); }
    static double log(double v) { return ::log(v); }
    static double asinh(double v) { return ::asinh(v); }
    static double abs(double v) { return ::fabs(v); }
    static double pi() { return M_PI; }
    static double half_pi() { return M_PI / 2; }  // core.jl:19,23
    static double below_one() { return ::nextafter(1.0, 0.0); }
    static double tiny() { return DBL_MIN; }
    static bool sq_overflows(double) { return false; }
};
template <> struct P<q128> {  // Double64 computed in binary128, range limits of Float64
    static q128 sinh(q128 v) { return ::sinhq(v); }
    static q128 cosh(q128 v) { return ::coshq(v); }
    static q128 tanh(q128 v) { return ::tanhq(v); }
    static q128 exp(q128 v) { return ::expq(v); }
    static q128 log(q128 v) { return ::logq(v); }
    static q128 asinh(q128 v) { return ::asinhq(v); }
    static q128 abs(q128 v) { return ::fabsq(v); }
    static q128 pi() { return M_PIq; }
    static q128 half_pi() { return M_PIq * (q128)0.5; }  // core.jl:25
    static q128 below_one() { return (q128)1 - ::ldexpq((q128)1, -106); }
    static q128 tiny() { return DBL_MIN; }
    static bool sq_overflows(q128 c2) { return c2 > (q128)DBL_MAX; }
};

// ---- L0: scalar primitives (core.jl:126-185) ---------------------------------------------------
template <class T> T node_x(T t)
{
    T v = P<T>::tanh(P<T>::half_pi() * P<T>::sinh(t));
    const T lim = P<T>::below_one();
    if (v >= (T)1) return lim;
    if (v <= -(T)1) return -lim;
    return v;
}
template <class T> T node_w(T t)
{
    const T hp = P<T>::half_pi();
    const T u = hp * P<T>::sinh(t);
    if (P<T>::abs(u) > (T)700.0) return (T)0;  // core.jl:151-153,179-181
    const T ch = P<T>::cosh(u);
    const T den = ch * ch;
    if (P<T>::sq_overflows(den)) return (T)0;
    return (hp * P<T>::cosh(t)) / den;
}
template <class T> T node_c(T t)  // 1-|x| without cancellation (core.jl:163-170)
{
    const T two = (T)1 + (T)1;
    const T u = P<T>::half_pi() * P<T>::sinh(P<T>::abs(t));
    const T z = P<T>::exp(-two * u);
    return two * z / ((T)1 + z);
}

// ---- windows (core.jl:187-230) ----------------------------------------------------------------
template <class T> T window_x()
{
    const T p = P<T>::below_one();
    return P<T>::asinh(P<T>::log(((T)1 + p) / ((T)1 - p)) / P<T>::pi());
}
template <class T> T window_w(int D)
{
    const T rhs = P<T>::log(P<T>::pi()) - P<T>::log(P<T>::tiny()) / (T)D;
    T t = P<T>::asinh(rhs / P<T>::pi());
    for (int it = 0; it < 10; ++it) {
        const T f = P<T>::pi() * P<T>::sinh(t) - t - rhs;
        const T df = P<T>::pi() * P<T>::cosh(t) - 1;
        t = t - f / df;
    }
    return t;
}
template <class T> T window(int which, int D)
{
    const T tx = window_x<T>();
    if (which == 1) return tx;
    const T tw = window_w<T>(D);
    if (which == 2) return tw;
    return tx < tw ? tx : tw;
}

// ---- output adaptors ------------------------------------------------------------------------
template <class T> struct Out {
    typedef T elem;
    static void put(void* base, long i, T v) { ((T*)base)[i] = v; }
};
template <> struct Out<q128> {
    static void put(void* base, long i, q128 v)
    {
        fts_dd* o = (fts_dd*)base + i;
        o->hi = (double)v;
        o->lo = (double)(v - (q128)o->hi);
    }
};

// ---- L1: tables ------------------------------------------------------------------------------
template <class T> void gen_fixed(int N, int D, void* x, void* w, void* h_out)
{
    const int n = N / 2;  // core.jl:257
    const T tm = window<T>(0, D);
    const T h = tm / (T)n;
    Out<T>::put(h_out, 0, h);
    for (int i = 0; i < n; ++i) {
        // range(h, tm, length=n)[i+1]: affine interpolation evaluated beyond working precision
        T t = h;
        if (n > 1) {
            const q128 step = ((q128)tm - (q128)h) / (q128)(n - 1);
            t = (i == n - 1) ? tm : (T)((q128)h + (q128)i * step);
        }
        Out<T>::put(x, i, node_x<T>(t));
        Out<T>::put(w, i, node_w<T>(t));
    }
}

template <class T>
void gen_cache1d(int L, int complement, void* tm_out, void* ix, void* iw, void* ic, void* xs, void* ws, void* cs)
{
    const T tm = complement ? window<T>(2, 1) : window<T>(0, 1);  // caches.jl:47-48
    const T half = (T)1 / ((T)1 + (T)1);
    T h = tm * half;
    Out<T>::put(tm_out, 0, tm);
    const T t0[2] = {h, h + h};
    for (int k = 0; k < 2; ++k) {
        Out<T>::put(ix, k, node_x<T>(t0[k]));
        Out<T>::put(iw, k, node_w<T>(t0[k]));
        Out<T>::put(ic, k, node_c<T>(t0[k]));
    }
    for (int level = 1; level <= L; ++level) {
        h *= half;
        const long n = 1L << level, off = n - 2;
        for (long i = 1; i <= n; ++i) {
            const T t = (T)(2 * i - 1) * h;
            Out<T>::put(xs, off + i - 1, node_x<T>(t));
            Out<T>::put(ws, off + i - 1, node_w<T>(t));
            Out<T>::put(cs, off + i - 1, node_c<T>(t));
        }
    }
}

template <class T> void gen_cache_tensor(int D, int L, void* tm_out, void* ix, void* iw, void* xs, void* ws)
{
    const T tm = window<T>(0, D);
    const T half = (T)1 / ((T)1 + (T)1);
    T h = tm * half;
    Out<T>::put(tm_out, 0, tm);
    const T t0[2] = {h, h + h};
    for (int k = 0; k < 2; ++k) {
        Out<T>::put(ix, k, node_x<T>(t0[k]));
        Out<T>::put(iw, k, node_w<T>(t0[k]));
    }
    for (int level = 1; level <= L; ++level) {
        h *= half;
        const long n = 2L << level, off = n - 4;
        for (long i = 1; i <= n; ++i) {
            const T t = (T)i * h;
            Out<T>::put(xs, off + i - 1, node_x<T>(t));
            Out<T>::put(ws, off + i - 1, node_w<T>(t));
        }
    }
}

// ---- optimal-spacing grids (core.jl:276-302) -------------------------------------------------
// principal branch of Lambert W for z >= 0 (the reference calls LambertW.jl): Halley iteration
template <class T> T lambert_w0(T z)
{
    if (!(z > (T)0)) return (T)0;
    T w = z < (T)3 ? P<T>::log((T)1 + z) * (T)0.75 : P<T>::log(z) - P<T>::log(P<T>::log(z));
    for (int it = 0; it < 60; ++it) {
        const T e = P<T>::exp(w);
        const T f = w * e - z;
        const T w1 = w + (T)1;
        const T step = f / (e * w1 - ((w + (T)2) * f) / (w1 + w1));
        w = w - step;
        if (P<T>::abs(step) <= P<T>::abs(w) * (T)1e-40) break;
    }
    return w;
}
// hopt(T, N, d) = (2/N) W(2 d N)   core.jl:289-291
template <class T> T h_opt(int N, double d) { return ((T)2 / (T)N) * lambert_w0<T>((T)(d + d) * (T)N); }
// hmax(T, N, D) = tmax(T,D) / (N div 2)   core.jl:293-302
template <class T> T h_max(int N, int D) { return window<T>(0, D) / (T)(N / 2); }
// number of points of the range h:h:tm (core.jl:285): k*h <= tm, with the half-ulp slack of
// Julia's float ranges
template <class T> long opt_count(T h, T tm)
{
    const q128 r = (q128)tm / (q128)h;
    long k = (long)r;
    if ((q128)(k + 1) - r < (q128)1e-9 * r) ++k;
    return k;
}
template <class T> long gen_opt(int N, double d, int D, long capacity, void* x, void* w, void* h_out)
{
    const T tm = window<T>(0, D);
    const T h = h_opt<T>(N, d);
    Out<T>::put(h_out, 0, h);
    if (!(h > (T)0)) return 0;
    const long k = opt_count<T>(h, tm);
    for (long i = 0; i < k && i < capacity; ++i) {
        const T t = (T)((q128)(i + 1) * (q128)h);  // range element in twice the working precision
        Out<T>::put(x, i, node_x<T>(t));
        Out<T>::put(w, i, node_w<T>(t));
    }
    return k;
}

}  // namespace

// entry points used by fts_host.cu (not part of the public ABI; the public wrappers validate)
extern "C" {

int fts_tg_hopt(int dtype, int N, double d, void* out)
{
    switch (dtype) {
    case FTS_F32: Out<float>::put(out, 0, h_opt<float>(N, d)); return 0;
    case FTS_F64: Out<double>::put(out, 0, h_opt<double>(N, d)); return 0;
    case FTS_DD64: Out<q128>::put(out, 0, h_opt<q128>(N, d)); return 0;
    }
    return 1;
}
int fts_tg_hmax(int dtype, int N, int D, void* out)
{
    switch (dtype) {
    case FTS_F32: Out<float>::put(out, 0, h_max<float>(N, D)); return 0;
    case FTS_F64: Out<double>::put(out, 0, h_max<double>(N, D)); return 0;
    case FTS_DD64: Out<q128>::put(out, 0, h_max<q128>(N, D)); return 0;
    }
    return 1;
}
long fts_tg_opt(int dtype, int N, double d, int D, long capacity, void* x, void* w, void* h)
{
    switch (dtype) {
    case FTS_F32: return gen_opt<float>(N, d, D, capacity, x, w, h);
    case FTS_F64: return gen_opt<double>(N, d, D, capacity, x, w, h);
    case FTS_DD64: return gen_opt<q128>(N, d, D, capacity, x, w, h);
    }
    return -1;
}

int fts_tg_window(int dtype, int which, int D, void* out)
{
    switch (dtype) {
    case FTS_F32: Out<float>::put(out, 0, window<float>(which, D)); return 0;
    case FTS_F64: Out<double>::put(out, 0, window<double>(which, D)); return 0;
    case FTS_DD64: Out<q128>::put(out, 0, window<q128>(which, D)); return 0;
    }
    return 1;
}
int fts_tg_fixed(int dtype, int N, int D, void* x, void* w, void* h)
{
    switch (dtype) {
    case FTS_F32: gen_fixed<float>(N, D, x, w, h); return 0;
    case FTS_F64: gen_fixed<double>(N, D, x, w, h); return 0;
    case FTS_DD64: gen_fixed<q128>(N, D, x, w, h); return 0;
    }
    return 1;
}
int fts_tg_cache1d(int dtype, int L, int complement, void* tm, void* ix, void* iw, void* ic, void* xs, void* ws, void* cs)
{
    switch (dtype) {
    case FTS_F32: gen_cache1d<float>(L, complement, tm, ix, iw, ic, xs, ws, cs); return 0;
    case FTS_F64: gen_cache1d<double>(L, complement, tm, ix, iw, ic, xs, ws, cs); return 0;
    case FTS_DD64: gen_cache1d<q128>(L, complement, tm, ix, iw, ic, xs, ws, cs); return 0;
    }
    return 1;
}
int fts_tg_cache_tensor(int dtype, int D, int L, void* tm, void* ix, void* iw, void* xs, void* ws)
{
    switch (dtype) {
    case FTS_F32: gen_cache_tensor<float>(D, L, tm, ix, iw, xs, ws); return 0;
    case FTS_F64: gen_cache_tensor<double>(D, L, tm, ix, iw, xs, ws); return 0;
    case FTS_DD64: gen_cache_tensor<q128>(D, L, tm, ix, iw, xs, ws); return 0;
    }
    return 1;
}
}  // extern "C"

// tensor table with complements on a caller-chosen window tm (2D complement extension):
// nodes t = i*h_l (caches.jl:173-177), c as caches.jl:84
template <class T> static void gen_tensor_cmpl(int L, T tm, T* ix, T* iw, T* ic, T* xs, T* ws, T* cs)
{
    const T half = (T)1 / ((T)1 + (T)1);
    T h = tm * half;
    const T t0[2] = {h, h + h};
    for (int k = 0; k < 2; ++k) {
        ix[k] = node_x<T>(t0[k]);
        iw[k] = node_w<T>(t0[k]);
        ic[k] = node_c<T>(t0[k]);
    }
    for (int level = 1; level <= L; ++level) {
        h *= half;
        const long n = 2L << level, off = n - 4;
        for (long i = 1; i <= n; ++i) {
            const T t = (T)i * h;
            xs[off + i - 1] = node_x<T>(t);
            ws[off + i - 1] = node_w<T>(t);
            cs[off + i - 1] = node_c<T>(t);
        }
    }
}
extern "C" {
int fts_tg_tensor_complements(int dtype, int L, const void* tm_in, void* ix, void* iw, void* ic, void* xs, void* ws, void* cs)
{
    if (dtype == FTS_F64) {
        gen_tensor_cmpl<double>(L, *(const double*)tm_in, (double*)ix, (double*)iw, (double*)ic, (double*)xs, (double*)ws, (double*)cs);
        return 0;
    }
    if (dtype == FTS_F32) {
        gen_tensor_cmpl<float>(L, *(const float*)tm_in, (float*)ix, (float*)iw, (float*)ic, (float*)xs, (float*)ws, (float*)cs);
        return 0;
    }
    return 1;
}
}
