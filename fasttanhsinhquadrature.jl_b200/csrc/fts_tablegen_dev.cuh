// fts_tablegen_dev.cuh — tanh-sinh node tables generated ON THE DEVICE.
//
// The same formulas, clamps and guards as the host generator (fts_tablegen.cpp) and the reference
// (/root/reference/src/core.jl:126-185 nodes, :256-267 fixed grid, src/caches.jl:59-91 1D levels,
// :148-188 tensor levels), evaluated by one thread per node with CUDA's libm in Float32/Float64
// and with double-double series/exp in Double64.  The tables are written straight into the
// interleaved Node / NodeC arrays the integration kernels read: no host arrays, no upload.
#pragma once
#include "fts_kernels.cuh"

namespace fts {

// ---- element-precision math of the generators ------------------------------------------------
template <class T> struct TG;
template <> struct TG<float> {
    static __device__ float sinh(float v) { return ::sinhf(v); }
    static __device__ float cosh(float v) { return ::coshf(v); }
    static __device__ float tanh(float v) { return ::tanhf(v); }
    static __device__ float exp(float v) { return ::expf(v); }
    static __device__ float below_one() { return 0.99999994f; }  // prevfloat(1f0)
    static __device__ bool overflowed(float v) { return !(v <= 3.402823466e+38f); }
    static __device__ float from_int(long long i) { return (float)i; }
    static __device__ float scale2(float v, int e) { return ::ldexpf(v, e); }
    // element i (0-based) of range(h, tm, length=n), interpolated beyond the working precision
    static __device__ float affine(float h, float tm, int i, int n) { return (float)((double)h + (double)i * (((double)tm - (double)h) / (double)(n - 1))); }
};
template <> struct TG<double> {
    static __device__ double sinh(double v) { return ::sinh(v); }
    static __device__ double cosh(double v) { return ::cosh(v); }
    static __device__ double tanh(double v) { return ::tanh(v); }
    static __device__ double exp(double v) { return ::exp(v); }
    static __device__ double below_one() { return 0.99999999999999989; }  // prevfloat(1.0)
    static __device__ bool overflowed(double v) { return !(v <= 1.7976931348623157e308); }
    static __device__ double from_int(long long i) { return (double)i; }
    static __device__ double scale2(double v, int e) { return ::ldexp(v, e); }
    static __device__ double affine(double h, double tm, int i, int n)
    {
        const dd step = (dd(tm) - dd(h)) / dd((double)(n - 1));
        return (dd(h) + step * (double)i).hi;
    }
};
template <> struct TG<dd> {
    // sinh for 0 <= v: Taylor series below 1/2 (no cancellation), (e - 1/e)/2 above
    static __device__ dd sinh(dd v)
    {
        if (v.hi < 0.5) {
            const dd v2 = v * v;
            dd term = v, s = v;
            for (int k = 1; k <= 18; ++k) {
                term = (term * v2) / dd((double)((2 * k) * (2 * k + 1)));
                s = s + term;
            }
            return s;
        }
        const dd e = dd_exp(v);
        return dd_mul_pwr2(e - dd(1.0) / e, 0.5);
    }
    static __device__ dd cosh(dd v)
    {
        const dd e = dd_exp(v);
        return dd_mul_pwr2(e + dd(1.0) / e, 0.5);
    }
    static __device__ dd tanh(dd v)  // 0 <= v
    {
        if (v.hi > 40.0) return dd(1.0);  // 1 - 2e^-80 rounds to 1 in 106 bits
        if (v.hi < 0.5) return sinh(v) / cosh(v);
        const dd z = dd_exp(-dd_mul_pwr2(v, 2.0));
        return (dd(1.0) - z) / (dd(1.0) + z);
    }
    static __device__ dd exp(dd v) { return dd_exp(v); }
    static __device__ dd below_one() { return dd(1.0, -1.232595164407830946e-32); }  // 1 - 2^-106
    static __device__ bool overflowed(dd v) { return !(v.hi <= 1.7976931348623157e308); }
    static __device__ dd from_int(long long i) { return dd((double)i); }
    static __device__ dd scale2(dd v, int e) { return dd_ldexp(v, e); }
    static __device__ dd affine(dd h, dd tm, int i, int n) { return h + ((tm - h) / dd((double)(n - 1))) * (double)i; }
};

// ordinate, weight, ordinate_complement of one node t > 0   (core.jl:126-185)
template <class T> __device__ void tg_node(T t, T& x, T& w, T& c)
{
    const T one = num<T>::one(), two = one + one;
    const T hp = num<T>::half_pi();
    const T u = hp * TG<T>::sinh(t);
    T v = TG<T>::tanh(u);
    const T lim = TG<T>::below_one();
    x = (v >= one) ? lim : v;  // core.jl:133-137 (t > 0: only the upper clamp can bind)
    if (num<T>::to_double(u) > 700.0) {
        w = num<T>::zero();  // core.jl:151-153, 179-181
    } else {
        const T ch = TG<T>::cosh(u);
        const T den = ch * ch;
        w = TG<T>::overflowed(den) ? num<T>::zero() : (hp * TG<T>::cosh(t)) / den;
    }
    const T z = TG<T>::exp(-(two * u));  // core.jl:163-170
    c = two * z / (one + z);
}

// fixed grid: node i of tanhsinh(T, N; D)   core.jl:256-267
template <class T> __global__ void k_tablegen_fixed(Node<T>* out, int n, T h, T tm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T t = h;
    if (n > 1) t = (i == n - 1) ? tm : TG<T>::affine(h, tm, i, n);
    T x, w, c;
    tg_node<T>(t, x, w, c);
    out[i].x = x;
    out[i].w = w;
}

// adaptive level tables.  tensor == 0: 1D cache, level l (1..L) holds the NEW nodes
// t = (2i-1) tm/2^(l+1), i = 1..2^l, at offset 2^l-2 (caches.jl:74-89); tensor != 0: level l holds
// ALL nodes t = i tm/2^(l+1), i = 1..2^(l+1), at offset 2^(l+1)-4 (caches.jl:168-186).
template <class T> __global__ void k_tablegen_levels(Node<T>* nodes, NodeC<T>* nodes_c, long long count, T tm, int tensor)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    int level;
    long long i;
    if (tensor) {
        const int top = 63 - __clzll(g + 4);  // = level + 1
        level = top - 1;
        i = g - ((1LL << top) - 4) + 1;
    } else {
        level = 63 - __clzll(g + 2);
        i = 2 * (g - ((1LL << level) - 2) + 1) - 1;
    }
    const T h = TG<T>::scale2(tm, -(level + 1));  // tm halved level+1 times: exact
    const T t = TG<T>::from_int(i) * h;
    T x, w, c;
    tg_node<T>(t, x, w, c);
    nodes[g].x = x;
    nodes[g].w = w;
    if (nodes_c) {
        nodes_c[g].x = x;
        nodes_c[g].w = w;
        nodes_c[g].c = c;
        nodes_c[g].pad = num<T>::zero();
    }
}

// the two level-0 nodes t = tm/2, tm (caches.jl:65-71): out = {x0, x1, w0, w1, c0, c1}
template <class T> __global__ void k_tablegen_initial(T* out, T tm)
{
    const int k = threadIdx.x;
    if (k >= 2) return;
    const T t = k == 0 ? TG<T>::scale2(tm, -1) : tm;
    T x, w, c;
    tg_node<T>(t, x, w, c);
    out[k] = x;
    out[2 + k] = w;
    out[4 + k] = c;
}

}  // namespace fts
