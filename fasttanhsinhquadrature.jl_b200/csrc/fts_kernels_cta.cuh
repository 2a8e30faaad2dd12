// fts_kernels_cta.cuh — cooperative (FTS_ORDER_FAST) kernels: the analogue of the reference's
// @turbo `_avx` twins (reassociated sums, FMA contraction), re-designed for the B200:
//
//   k_*_cta    one CTA per integral.  The coordinates x0±Δx·x_i (and y, z, weights, complement
//              distances) of the current level are staged in SHARED memory by the whole CTA —
//              the reference hoists them into per-cache scratch vectors that live in the globally
//              memoised cache (adaptive.jl:570-580, a latent race); here the scratch is per CTA.
//              The level's NEW cells (index-parity classes of adaptive.jl:199-209, 430-456) are
//              enumerated by a flat index and dealt to threads with stride blockDim, so every
//              thread gets the same number of cells (±1) whatever the class mix.
//   k_level*_grid  one launch per refinement level, many CTAs per integral, for ONE large 2D/3D
//              integral; the flat cell index is additionally dealt over `nranks` devices.  CTA
//              partials are merged by the last CTA (ticket) in a fixed order -> deterministic.
//
// Sums: plain FMA inside a thread's short run, TwoSum-compensated (s, c) pairs for every merge
// (warp shuffle -> shared memory -> CTA -> grid -> ranks) and for the running total across levels.
#pragma once
#include "fts_kernels.cuh"

namespace fts {

constexpr int CTA_THREADS = 256;

template <class T> __device__ __forceinline__ Comp<T> warp_reduce(Comp<T> v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Comp<T> o;
        o.s = __shfl_down_sync(0xffffffffu, v.s, off);
        o.c = __shfl_down_sync(0xffffffffu, v.c, off);
        v.merge(o);
    }
    return v;
}

// result valid in thread 0; `red` is shared memory for 64 elements; contains two barriers
template <class T> __device__ __forceinline__ Comp<T> block_reduce(Comp<T> v, T* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_reduce(v);
    if (lane == 0) {
        red[2 * warp] = v.s;
        red[2 * warp + 1] = v.c;
    }
    __syncthreads();
    if (warp == 0) {
        Comp<T> w;
        if (lane < nw) {
            w.s = red[2 * lane];
            w.c = red[2 * lane + 1];
        }
        v = warp_reduce(w);
    }
    __syncthreads();
    return v;
}

template <class T> __device__ __forceinline__ Comp<T> comp_of(T v)
{
    Comp<T> c;
    c.s = v;
    return c;
}

// ------------------------------------------------------------------------------------ 1D -----
// adaptive_integrate_1D_avx / adaptive_integrate_1D_cmpl_avx   adaptive.jl:94-133, 805-855
template <class T, class F, bool CMPL> __device__ __forceinline__ void adaptive1d_cta_body(const Args<T>& a)
{
    family_tables_init<T, F>();
    __shared__ T red[64];
    __shared__ int s_done;
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half(), one = num<T>::one();
    T lo = a.low[b * a.bstride], hi = a.up[b * a.bstride];
    bool neg = false;
    int flags = 0;
    if (a.options & OPT_QUAD_EDGE) {
        if (!quad_edge_axis(lo, hi, neg)) {
            if (tid == 0) store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
            return;
        }
        if (neg) flags |= FLAG_FLIPPED;
    }
    const T dx = half * (hi - lo), x0 = half * (hi + lo);
    T h = a.tm * half;
    // level 0: centre + 2 node pairs on threads 0..2
    T acc = num<T>::zero();
    if (tid == 0) {
        if constexpr (CMPL) acc = num<T>::half_pi() * F::template eval<T, true>(x0, dx, dx, p);
        else acc = num<T>::half_pi() * F::template eval<T, true>(x0, p);
    } else if (tid <= 2) {
        const T xk = a.ix[tid - 1], wk = a.iw[tid - 1];
        const T d = dx * xk;
        if constexpr (CMPL) {
            const T dxck = dx * a.ic[tid - 1], dx1p = dx * (one + xk);
            acc = wk * (F::template eval<T, true>(d + x0, dxck, dx1p, p) + F::template eval<T, true>(x0 - d, dx1p, dxck, p));
        } else {
            acc = wk * (F::template eval<T, true>(d + x0, p) + F::template eval<T, true>(x0 - d, p));
        }
    }
    Comp<T> s_total = block_reduce(comp_of(acc), red);
    T old_res = dx * h * s_total.value();
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        const int n = 1 << level;
        acc = num<T>::zero();
        if constexpr (CMPL) {
            const NodeC<T>* __restrict__ lv = a.nodes_c + (n - 2);
            for (int i = tid; i < n; i += blockDim.x) {
                const NodeC<T> nd = lv[i];
                const T d = dx * nd.x, dxck = dx * nd.c, dx1p = dx * (one + nd.x);
                acc = fm::fma(nd.w, F::template eval<T, true>(d + x0, dxck, dx1p, p) + F::template eval<T, true>(x0 - d, dx1p, dxck, p), acc);
            }
        } else {
            const Node<T>* __restrict__ lv = a.nodes + (n - 2);
            for (int i = tid; i < n; i += blockDim.x) {
                const Node<T> nd = ldg_node(lv + i);
                const T xp = fm::fma(dx, nd.x, x0), xm = fm::fma(-dx, nd.x, x0);
                acc = fm::fma(nd.w, F::template eval<T, true>(xp, p) + F::template eval<T, true>(xm, p), acc);
            }
        }
        const Comp<T> lv_sum = block_reduce(comp_of(acc), red);
        if (tid == 0) {
            s_total.merge(lv_sum);
            const T new_res = dx * h * s_total.value();
            err = fm::abs(new_res - old_res);
            old_res = new_res;
            level_used = level;
            const bool conv = err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels);
            if (conv) flags |= FLAG_CONVERGED;
            s_done = conv ? 1 : 0;
        }
        __syncthreads();
        if (s_done) break;
    }
    if (tid == 0) {
        if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
}

template <class T, class F> __global__ void __launch_bounds__(CTA_THREADS) k_adaptive1d_cta(const Args<T> a) { adaptive1d_cta_body<T, F, false>(a); }
template <class T, class F> __global__ void __launch_bounds__(CTA_THREADS) k_adaptive1d_cmpl_cta(const Args<T> a) { adaptive1d_cta_body<T, F, true>(a); }

// integrate1D_avx   integrate1D.jl:138-147
template <class T, class F> __global__ void __launch_bounds__(CTA_THREADS) k_fixed1d_cta(const Args<T> a)
{
    family_tables_init<T, F>();
    __shared__ T red[64];
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T lo = a.low[b * a.bstride], hi = a.up[b * a.bstride];
    const T dx = half * (hi - lo), x0 = half * (hi + lo);
    T acc = tid == 0 ? num<T>::half_pi() * F::template eval<T, true>(x0, p) : num<T>::zero();
    for (int i = tid; i < a.n; i += blockDim.x) {
        const Node<T> nd = ldg_node(a.grid + i);
        const T xp = fm::fma(dx, nd.x, x0), xm = fm::fma(-dx, nd.x, x0);
        acc = fm::fma(nd.w, F::template eval<T, true>(xm, p) + F::template eval<T, true>(xp, p), acc);
    }
    const Comp<T> s = block_reduce(comp_of(acc), red);
    if (tid == 0) a.value[b] = dx * a.h * s.value();
}

// ------------------------------------------------------------------------------------ 2D -----
template <class T> struct Sm2 {
    T *xp, *xm, *yp, *ym, *w;   // hoisted coordinates and weights of the level
    T *bx, *ax, *dyc, *cy;      // complement distances Δx·c, Δx·(1+x), Δy·c, Δy·(1+x)
    T dx, x0, dy, y0, w0;
};

template <class T, bool CMPL> __device__ __forceinline__ size_t sm2_elems(int n) { return (size_t)(CMPL ? 9 : 5) * n; }

template <class T, bool CMPL> __device__ __forceinline__ void sm2_carve(Sm2<T>& s, T* base, int n)
{
    s.xp = base; s.xm = base + n; s.yp = base + 2 * n; s.ym = base + 3 * n; s.w = base + 4 * n;
    if (CMPL) { s.bx = base + 5 * n; s.ax = base + 6 * n; s.dyc = base + 7 * n; s.cy = base + 8 * n; }
}

// stage one level; `x,w,c` given as strided element pointers so that Node / NodeC / initial
// tuples all fit (stride in elements)
template <class T, bool CMPL>
__device__ __forceinline__ void sm2_stage(Sm2<T>& s, const T* x, const T* w, const T* c, int stride, int n, int tid = threadIdx.x,
                                          int nthr = blockDim.x)
{
    const T one = num<T>::one();
    for (int i = tid; i < n; i += nthr) {
        const T xi = x[(size_t)i * stride], wi = w[(size_t)i * stride];
        s.xp[i] = fm::fma(s.dx, xi, s.x0);
        s.xm[i] = fm::fma(-s.dx, xi, s.x0);
        s.yp[i] = fm::fma(s.dy, xi, s.y0);
        s.ym[i] = fm::fma(-s.dy, xi, s.y0);
        s.w[i] = wi;
        if (CMPL) {
            const T ci = c[(size_t)i * stride];
            s.bx[i] = s.dx * ci;
            s.ax[i] = s.dx * (one + xi);
            s.dyc[i] = s.dy * ci;
            s.cy[i] = s.dy * (one + xi);
        }
    }
}

template <class T, class F, bool CMPL> __device__ __forceinline__ T cell2(const Sm2<T>& s, const T* p, int i, int j)
{
    const T xp = s.xp[i], xm = s.xm[i], yp = s.yp[j], ym = s.ym[j];
    T f;
    if constexpr (CMPL) {
        const T bx = s.bx[i], ax = s.ax[i], dyc = s.dyc[j], cy = s.cy[j];
        // the four reflections side by side (families with a batch evaluator interleave their chains)
        const T xs[4] = {xp, xm, xp, xm}, ys[4] = {yp, yp, ym, ym};
        const T b_[4] = {bx, ax, bx, ax}, a_[4] = {ax, bx, ax, bx}, d_[4] = {dyc, dyc, cy, cy}, c_[4] = {cy, cy, dyc, dyc};
        T v[4];
        eval2dc_n<T, F, true, 4>(xs, ys, b_, a_, d_, c_, p, v);
        f = (v[0] + v[1]) + (v[2] + v[3]);
    } else if constexpr (HasEvalN<F>::value) {
        const T xs[4] = {xp, xm, xp, xm}, ys[4] = {yp, yp, ym, ym};
        T v[4];
        eval2d_n<T, F, true, 4, UsesExpTab<T, F>::value ? EXP_SMEM : 0>(xs, ys, p, v);
        f = (v[0] + v[1]) + (v[2] + v[3]);
    } else {
        f = (F::template eval<T, true>(xp, yp, p) + F::template eval<T, true>(xm, yp, p)) +
            (F::template eval<T, true>(xp, ym, p) + F::template eval<T, true>(xm, ym, p));
    }
    return (s.w[i] * s.w[j]) * f;
}

template <class T, class F, bool CMPL> __device__ __forceinline__ T axis2(const Sm2<T>& s, const T* p, int i)
{
    T f;
    if constexpr (CMPL) {
        // the four evaluations side by side (same values as four F::eval calls: the batch evaluators are bit-identical)
        const T xs[4] = {s.xp[i], s.xm[i], s.x0, s.x0}, ys[4] = {s.y0, s.y0, s.yp[i], s.ym[i]};
        const T b_[4] = {s.bx[i], s.ax[i], s.dx, s.dx}, a_[4] = {s.ax[i], s.bx[i], s.dx, s.dx};
        const T d_[4] = {s.dy, s.dy, s.dyc[i], s.cy[i]}, c_[4] = {s.dy, s.dy, s.cy[i], s.dyc[i]};
        T v[4];
        eval2dc_n<T, F, true, 4>(xs, ys, b_, a_, d_, c_, p, v);
        f = (v[0] + v[1]) + (v[2] + v[3]);
    } else {
        f = (F::template eval<T, true>(s.xp[i], s.y0, p) + F::template eval<T, true>(s.xm[i], s.y0, p)) +
            (F::template eval<T, true>(s.x0, s.yp[i], p) + F::template eval<T, true>(s.x0, s.ym[i], p));
    }
    return (s.w[i] * s.w0) * f;
}

// Sum of the cells a thread owns.  all_new: every (i,j) (level 0 / fixed grids, n arbitrary);
// otherwise n = 2^log2n and only cells with an odd 1-based index in i or j are new
// (adaptive.jl:199-209): 0-based i even -> all j; i odd -> j even.
template <class T, class F, bool CMPL>
__device__ __forceinline__ T level_sum_2d(const Sm2<T>& s, const T* p, int n, int log2n, bool all_new, long long start, long long stride)
{
    T acc = num<T>::zero();
    if (all_new) {
        const long long ncell = (long long)n * n;
        for (long long t = start; t < ncell; t += stride) {
            const int i = (int)(t / n), j = (int)(t - (long long)i * n);
            acc = acc + cell2<T, F, CMPL>(s, p, i, j);
        }
        for (long long t = start; t < n; t += stride) acc = acc + axis2<T, F, CMPL>(s, p, (int)t);
    } else {
        const long long nA = ((long long)n * n) >> 1, ncell = nA + (nA >> 1);
        for (long long t = start; t < ncell; t += stride) {
            int i, j;
            if (t < nA) {
                i = (int)(t >> log2n) << 1;
                j = (int)(t & (n - 1));
            } else {
                const long long u = t - nA;
                i = ((int)(u >> (log2n - 1)) << 1) | 1;
                j = (int)(u & ((n >> 1) - 1)) << 1;
            }
            acc = acc + cell2<T, F, CMPL>(s, p, i, j);
        }
        for (long long t = start; t < (n >> 1); t += stride) acc = acc + axis2<T, F, CMPL>(s, p, (int)t << 1);
    }
    return acc;
}

// ---- warp-shaped level sum: a lane owns a column ------------------------------------------------------
// x-side arrays of the level through 32-bit shared-window addresses that the compiler must keep in
// registers (`asm("" : "+r")`): left to itself it re-derives the nine array pointers from threadIdx and the
// launch arguments in every trip of the cell loop (35 of 91 instructions per cell) rather than spend
// registers under the 72-register bound that keeps 28 warps on an SM.
template <class T> struct XSide { unsigned xp, xm, w, bx, ax; };
__device__ __forceinline__ unsigned smem_addr_pinned(const void* p)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm("" : "+r"(a));
    return a;
}
__device__ __forceinline__ double lds_elem(unsigned addr, double)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds_elem(unsigned addr, float)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
template <class T, bool CMPL> __device__ __forceinline__ XSide<T> make_xside(const Sm2<T>& s)
{
    XSide<T> x;
    x.xp = smem_addr_pinned(s.xp); x.xm = smem_addr_pinned(s.xm); x.w = smem_addr_pinned(s.w);
    x.bx = CMPL ? smem_addr_pinned(s.bx) : 0u;
    x.ax = CMPL ? smem_addr_pinned(s.ax) : 0u;
    return x;
}
// One cell (i, j) of a column whose y-side values the lane keeps in registers: the 4 reflections are evaluated
// side by side.  (Seven warps per scheduler are resident in the warp-shaped kernel: 4 chains per warp already
// fill the FP64 pipe, and a second cell per trip costs the registers that keep 28 warps on an SM.)
template <class T> struct Col2 { T yp, ym, dyc, cy, wj; };
template <class T, class F, bool CMPL> __device__ __forceinline__ T cell_col(const XSide<T>& x, const T* p, const Col2<T>& c, unsigned off)
{
    T v[4];
    if constexpr (CMPL) {
        // complement families are functions of the distances: the plain coordinates are only fetched when the body reads them
        const T bx = lds_elem(x.bx + off, T()), ax = lds_elem(x.ax + off, T());
        const T xp = lds_elem(x.xp + off, T()), xm = lds_elem(x.xm + off, T());
        const T xs[4] = {xp, xm, xp, xm}, ys[4] = {c.yp, c.yp, c.ym, c.ym};
        const T b_[4] = {bx, ax, bx, ax}, a_[4] = {ax, bx, ax, bx}, d_[4] = {c.dyc, c.dyc, c.cy, c.cy}, c_[4] = {c.cy, c.cy, c.dyc, c.dyc};
        eval2dc_n<T, F, true, 4>(xs, ys, b_, a_, d_, c_, p, v);
    } else {
        const T xp = lds_elem(x.xp + off, T()), xm = lds_elem(x.xm + off, T());
        const T xs[4] = {xp, xm, xp, xm}, ys[4] = {c.yp, c.yp, c.ym, c.ym};
        eval2d_n<T, F, true, 4, UsesExpTab<T, F>::value ? EXP_SMEM : 0>(xs, ys, p, v);
    }
    return (lds_elem(x.w + off, T()) * c.wj) * ((v[0] + v[1]) + (v[2] + v[3]));
}
template <class T, bool CMPL> __device__ __forceinline__ Col2<T> load_col(const Sm2<T>& s, int j)
{
    Col2<T> c;
    c.yp = s.yp[j]; c.ym = s.ym[j]; c.wj = s.w[j];
    if constexpr (CMPL) { c.dyc = s.dyc[j]; c.cy = s.cy[j]; }
    else { c.dyc = num<T>::zero(); c.cy = num<T>::zero(); }
    return c;
}
// New cells of a level for ONE WARP (0-based: rows i even -> every column, rows i odd -> even columns), n a
// power of two >= 4: a lane owns a column j, keeps its y-side values in registers and walks down the rows —
// the lanes of a row group read the same x-side entry (shared-memory broadcast), no flat-index decode per
// cell.  Levels with fewer than 32 columns split the rows over 32/columns lane groups.
template <class T, class F, bool CMPL, int N = 0> __device__ __forceinline__ T level_sum_2d_warp(const Sm2<T>& s, const T* p, int n_runtime, int lane)
{
    const int n = N > 0 ? N : n_runtime;  // N > 0: the column/row-group arithmetic below folds into constants
    const XSide<T> x = make_xside<T, CMPL>(s);
    const unsigned nbytes = (unsigned)n * (unsigned)sizeof(T);
    T acc = num<T>::zero();
    {   // even rows x all columns
        const int cpl = n < 32 ? n : 32, sh = __ffs(cpl) - 1, g = 32 >> sh;  // columns per pass (a power of two), row groups
        const unsigned off0 = (unsigned)(2 * (lane >> sh)) * (unsigned)sizeof(T), step = (unsigned)(2 * g) * (unsigned)sizeof(T);
        for (int j = lane & (cpl - 1); j < n; j += 32) {
            const Col2<T> c = load_col<T, CMPL>(s, j);
#pragma unroll 1
            for (unsigned off = off0; off < nbytes; off += step) acc = acc + cell_col<T, F, CMPL>(x, p, c, off);
        }
    }
    {   // odd rows x even columns
        const int half = n >> 1, cpl = half < 32 ? half : 32, sh = __ffs(cpl) - 1, g = 32 >> sh;
        const unsigned off0 = (unsigned)(1 + 2 * (lane >> sh)) * (unsigned)sizeof(T), step = (unsigned)(2 * g) * (unsigned)sizeof(T);
        for (int j = (lane & (cpl - 1)) << 1; j < n; j += 64) {
            const Col2<T> c = load_col<T, CMPL>(s, j);
#pragma unroll 1
            for (unsigned off = off0; off < nbytes; off += step) acc = acc + cell_col<T, F, CMPL>(x, p, c, off);
        }
    }
    for (int t = lane; t < (n >> 1); t += 32) acc = acc + axis2<T, F, CMPL>(s, p, t << 1);
    return acc;
}

// Four sums over the warp for the price of six shuffles: the lanes trade halves (xor 16, xor 8) before the plain
// butterfly, so that every lane of the 8-lane class m = lane >> 3 ends up with the warp's total of S[m].
template <class T> __device__ __forceinline__ T warp_reduce4(const T (&S)[4], int lane)
{
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
    T a0 = b4 ? S[2] : S[0], a1 = b4 ? S[3] : S[1];
    const T o0 = b4 ? S[0] : S[2], o1 = b4 ? S[1] : S[3];
    a0 = a0 + __shfl_xor_sync(0xffffffffu, o0, 16);
    a1 = a1 + __shfl_xor_sync(0xffffffffu, o1, 16);
    T k = b3 ? a1 : a0;
    const T o = b3 ? a0 : a1;
    k = k + __shfl_xor_sync(0xffffffffu, o, 8);
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) k = k + __shfl_xor_sync(0xffffffffu, k, off);
    return k;
}

extern __shared__ __align__(16) unsigned char fts_dyn_smem[];

// ---- bulk-async staging of a level's node slice (TMA unit: cp.async.bulk + mbarrier) -------------
// The node slice of level l+1 ({x,w} or {x,w,c,pad} records, contiguous in the cache) is copied
// global -> shared by the TMA unit WHILE the group evaluates level l: one elected thread posts the
// expected byte count on an mbarrier and issues the copy, nobody waits for global memory inside the
// level loop.  Two raw buffers and two barriers per group, used alternately; `use[b]` counts the
// completed phases of barrier b (the mbarrier's phase bit is its parity).
struct NodePipe {
    unsigned bar[2];   // shared-window addresses of the two mbarriers
    unsigned raw[2];   // shared-window addresses of the two raw slice buffers
    unsigned use[2];   // waits done on each barrier (uniform over the group's threads)
};
__device__ __forceinline__ void pipe_init(NodePipe& np, void* bars, void* raw0, void* raw1, bool elected)
{
    np.bar[0] = (unsigned)__cvta_generic_to_shared(bars);
    np.bar[1] = np.bar[0] + 8u;
    np.raw[0] = (unsigned)__cvta_generic_to_shared(raw0);
    np.raw[1] = (unsigned)__cvta_generic_to_shared(raw1);
    np.use[0] = np.use[1] = 0u;
    if (elected) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(np.bar[0]) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(np.bar[1]) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
// elected thread: post `bytes` on barrier b and start the copy of [src, src + bytes) into raw buffer b
__device__ __forceinline__ void pipe_issue(const NodePipe& np, int b, const void* src, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(np.bar[b]), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(np.raw[b]), "l"(src), "r"(bytes),
                 "r"(np.bar[b])
                 : "memory");
}
// every thread of the group: wait until the copy posted on barrier b has landed
__device__ __forceinline__ void pipe_wait(NodePipe& np, int b)
{
    const unsigned parity = np.use[b] & 1u;
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(np.bar[b]), "r"(parity) : "memory");
    } while (!done);
    np.use[b] += 1u;
}
template <class T, bool CMPL> __device__ __forceinline__ unsigned node_bytes() { return CMPL ? (unsigned)sizeof(NodeC<T>) : (unsigned)sizeof(Node<T>); }
// shared memory of one group of the 2D adaptive kernel, in bytes: derived coordinate arrays + two raw slices + two barriers
template <class T, bool CMPL> __host__ __device__ __forceinline__ size_t sm2_group_bytes(int ncap)
{
    return (size_t)(CMPL ? 9 : 5) * ncap * sizeof(T) + 2 * (size_t)ncap * (CMPL ? sizeof(NodeC<T>) : sizeof(Node<T>)) + 16;
}

// A "group" of G threads owns one integral: the whole CTA (G = CTA_THREADS) or one warp (G = 32).
template <int G> __device__ __forceinline__ void group_sync()
{
    if constexpr (G == 32) __syncwarp();
    else __syncthreads();
}
template <class T, int G> __device__ __forceinline__ Comp<T> group_reduce(Comp<T> v, T* red)  // valid in the group's thread 0
{
    if constexpr (G == 32) {
        // one warp = 32 partial sums: a plain shuffle tree (5 roundings, ~5e-16) instead of the TwoSum merges of
        // the CTA and grid reductions, which cost as many instructions per level as 5 cells; the running total
        // over the levels stays compensated
        T t = v.s + v.c;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t = t + __shfl_down_sync(0xffffffffu, t, off);
        return comp_of(t);
    } else {
        return block_reduce(v, red);
    }
}

// adaptive_integrate_2D_avx (+ tensor-complement extension)   adaptive.jl:242-333
// One integral by a group of G threads (`gtid` = index in the group).  `level_cap` < max_levels:
// an integral that has not converged after level_cap is queued (a.tail_idx) for a second pass by
// whole CTAs instead of going deeper here (the group's shared-memory slice holds 2^(level_cap+1)
// nodes per array).
template <class T, class F, int G>
__device__ __forceinline__ void adaptive2d_group(const Args<T>& a, long long b, int gtid, T* smem, int ncap, int level_cap, T* red, int* s_done, NodePipe& np)
{
    constexpr bool CMPL = is_cmpl2d<F>::value;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    T lo[2], hi[2];
    bool neg;
    int flags;
    if (!load_box<2>(a, b, lo, hi, neg, flags)) {
        if (gtid == 0) store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
        return;
    }
    const T half = num<T>::half();
    Sm2<T> s;
    s.dx = half * (hi[0] - lo[0]); s.x0 = half * (hi[0] + lo[0]);
    s.dy = half * (hi[1] - lo[1]); s.y0 = half * (hi[1] + lo[1]);
    s.w0 = num<T>::half_pi();
    sm2_carve<T, CMPL>(s, smem, ncap);
    const unsigned char* raw_base = reinterpret_cast<const unsigned char*>(smem + sm2_elems<T, CMPL>(ncap));
    const int deepest = a.max_levels < level_cap ? a.max_levels : level_cap;  // deepest level this pass can reach
    // node slice of `level` (n = 2^(level+1) records at offset n - 4) -> raw buffer level & 1, by the TMA unit
    auto prefetch = [&](int level) {
        if (gtid != 0 || level > deepest) return;
        const int n = 2 << level;
        const void* src = CMPL ? (const void*)(a.nodes_c + (n - 4)) : (const void*)(a.nodes + (n - 4));
        pipe_issue(np, level & 1, src, (unsigned)n * node_bytes<T, CMPL>());
    };
    prefetch(1);  // in flight while level 0 is evaluated
    T h = a.tm * half;
    // level 0: base grid {h, 2h}, all cells new (adaptive.jl:264-272)
    sm2_stage<T, CMPL>(s, a.ix, a.iw, a.ic, 1, 2, gtid, G);
    group_sync<G>();
    T acc = level_sum_2d<T, F, CMPL>(s, p, 2, 1, true, gtid, G);
    if (gtid == 0) {
        if constexpr (CMPL) acc = acc + (s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, s.dx, s.dx, s.dy, s.dy, p);
        else acc = acc + (s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, p);
    }
    Comp<T> s_total = group_reduce<T, G>(comp_of(acc), red);
    T old_res = s.dx * s.dy * (h * h) * s_total.value();
    T err = num<T>::zero();
    int level_used = 0;
    int in_flight = deepest >= 1 ? 1 : 0;  // level whose slice has been requested but not yet consumed (0: none)
    bool queued = false;
    for (int level = 1; level <= a.max_levels; ++level) {
        if (level > level_cap) {  // too deep for this group's shared-memory slice: whole-CTA pass
            if (gtid == 0) a.tail_idx[atomicAdd(a.tail_count, 1)] = (int)b;
            queued = true;
            break;
        }
        h = h * half;
        const int n = 2 << level;
        group_sync<G>();  // every thread is done with the previous level's coordinates
        pipe_wait(np, level & 1);  // the slice of this level has landed in shared memory
        in_flight = 0;
        if constexpr (CMPL) {
            const NodeC<T>* lv = reinterpret_cast<const NodeC<T>*>(raw_base + (size_t)(level & 1) * ncap * sizeof(NodeC<T>));
            sm2_stage<T, CMPL>(s, &lv->x, &lv->w, &lv->c, 4, n, gtid, G);
        } else {
            const Node<T>* lv = reinterpret_cast<const Node<T>*>(raw_base + (size_t)(level & 1) * ncap * sizeof(Node<T>));
            sm2_stage<T, CMPL>(s, &lv->x, &lv->w, nullptr, 2, n, gtid, G);
        }
        group_sync<G>();
        if (level + 1 <= deepest) {  // the other raw buffer was last read two levels ago: refill it during this level's evaluation
            prefetch(level + 1);
            in_flight = level + 1;
        }
        if (G == 32) acc = level_sum_2d_warp<T, F, CMPL>(s, p, n, gtid);
        else acc = level_sum_2d<T, F, CMPL>(s, p, n, level + 1, false, gtid, G);
        const Comp<T> lv_sum = group_reduce<T, G>(comp_of(acc), red);
        int done = 0;
        if (gtid == 0) {
            s_total.merge(lv_sum);
            const T new_res = s.dx * s.dy * (h * h) * s_total.value();
            err = fm::abs(new_res - old_res);
            old_res = new_res;
            level_used = level;
            const bool conv = err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels);
            if (conv) flags |= FLAG_CONVERGED;
            done = conv ? 1 : 0;
        }
        if constexpr (G == 32) {
            done = __shfl_sync(0xffffffffu, done, 0);
        } else {
            if (gtid == 0) *s_done = done;
            __syncthreads();
            done = *s_done;
        }
        if (done) break;
    }
    // a slice requested for a level that is never evaluated (converged, or handed to the queue) is still
    // consumed: the barrier's phase count stays in step for the group's next integral
    if (in_flight) pipe_wait(np, in_flight & 1);
    if (gtid == 0 && !queued) {
        if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
}

// ---- the warp-shaped 2D adaptive path ---------------------------------------------------------------------------
// One warp per integral, shallow levels (k_adaptive2d_warp).  Shared memory of a warp, NC = ncap + 1 entries
// (the last entry is the CENTRE node, see warp_stage_centre):
//   x-side RECORDS  [Δx·c, Δx·(1+x), x0+Δx·x, x0-Δx·x, w, pad]: the lanes that walk down a row read one record — a
//                   broadcast, two or three vector loads off ONE address register, no per-array address arithmetic
//   y-side ARRAYS   y0+Δy·x | y0-Δy·x | Δy·c | Δy·(1+x) | w: read once per column into registers
//   two raw node slices (TMA destinations) and their two mbarriers.
template <class T> struct WarpRec { static constexpr int ELEMS = sizeof(T) == 8 ? 6 : 8; };  // 48 / 32 bytes: 16-byte aligned records
template <class T, bool CMPL> __host__ __device__ __forceinline__ size_t warp2d_bytes(int ncap)
{
    const size_t nc = (size_t)ncap + 1;
    const size_t xy = nc * WarpRec<T>::ELEMS * sizeof(T) + ((5 * nc * sizeof(T) + 15) & ~(size_t)15);
    const size_t fallback = (size_t)(CMPL ? 9 : 5) * ncap * sizeof(T);  // adaptive2d_group's arrays (batches with max_levels < 4)
    return (xy > fallback ? xy : fallback) + 2 * (size_t)ncap * (CMPL ? sizeof(NodeC<T>) : sizeof(Node<T>)) + 16;
}
template <class T> struct WarpSm {
    T* xrec;        // records
    T* ycol;        // arrays, NC apart
    int nc;
    T dx, x0, dy, y0, w0;
};
__device__ __forceinline__ void lds_rec(unsigned addr, double& bx, double& ax, double& xp, double& xm)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(bx), "=d"(ax) : "r"(addr) : "memory");
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(xp), "=d"(xm) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lds_rec(unsigned addr, float& bx, float& ax, float& xp, float& xm)
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bx), "=f"(ax), "=f"(xp), "=f"(xm) : "r"(addr) : "memory");
}
__device__ __forceinline__ double lds_rec_w(unsigned addr, double)
{
    double w;
    asm volatile("ld.shared.f64 %0, [%1+32];" : "=d"(w) : "r"(addr) : "memory");
    return w;
}
__device__ __forceinline__ float lds_rec_w(unsigned addr, float)
{
    float w;
    asm volatile("ld.shared.f32 %0, [%1+16];" : "=f"(w) : "r"(addr) : "memory");
    return w;
}
// One node per lane: the hoist of adaptive.jl:570-580 (same expressions as sm2_stage) from a raw slice into
// the records and arrays.
template <class T, bool CMPL> __device__ __forceinline__ void warp_stage(const WarpSm<T>& s, const unsigned char* raw, int n, int lane)
{
    constexpr int RE = WarpRec<T>::ELEMS;
    const T one = num<T>::one();
    for (int i = lane; i < n; i += 32) {
        T xi, wi, ci = one;
        if constexpr (CMPL) {
            const NodeC<T> nd = reinterpret_cast<const NodeC<T>*>(raw)[i];
            xi = nd.x; wi = nd.w; ci = nd.c;
        } else {
            const Node<T> nd = reinterpret_cast<const Node<T>*>(raw)[i];
            xi = nd.x; wi = nd.w;
        }
        T* r = s.xrec + (size_t)i * RE;
        T* y = s.ycol + i;
        r[2] = fm::fma(s.dx, xi, s.x0);
        r[3] = fm::fma(-s.dx, xi, s.x0);
        r[4] = wi;
        y[0] = fm::fma(s.dy, xi, s.y0);
        y[s.nc] = fm::fma(-s.dy, xi, s.y0);
        y[4 * s.nc] = wi;
        if constexpr (CMPL) {
            r[0] = s.dx * ci;
            r[1] = s.dx * (one + xi);
            y[2 * s.nc] = s.dy * ci;
            y[3 * s.nc] = s.dy * (one + xi);
        }
    }
}
// Entry ncap (the last one, whatever the level) is the CENTRE node with HALF its weight, staged once per integral: a
// cell on the centre row or column has two pairs of coincident reflections, which the cell evaluator counts twice —
// w_i (w0/2) (2 f + 2 f') is exactly the axis term w_i w0 (f + f') of adaptive.jl:204-206.
template <class T, bool CMPL> __device__ __forceinline__ void warp_stage_centre(const WarpSm<T>& s)
{
    const int i = s.nc - 1;
    T* r = s.xrec + (size_t)i * WarpRec<T>::ELEMS;
    T* y = s.ycol + i;
    const T wc = num<T>::half() * s.w0;
    r[2] = s.x0; r[3] = s.x0; r[4] = wc;
    y[0] = s.y0; y[s.nc] = s.y0; y[4 * s.nc] = wc;
    if constexpr (CMPL) { r[0] = s.dx; r[1] = s.dx; y[2 * s.nc] = s.dy; y[3 * s.nc] = s.dy; }
}
template <class T, bool CMPL> __device__ __forceinline__ Col2<T> warp_col(const WarpSm<T>& s, int j)
{
    Col2<T> c;
    const T* y = s.ycol + j;
    c.yp = y[0]; c.ym = y[s.nc]; c.wj = y[4 * s.nc];
    if constexpr (CMPL) { c.dyc = y[2 * s.nc]; c.cy = y[3 * s.nc]; }
    else { c.dyc = num<T>::zero(); c.cy = num<T>::zero(); }
    return c;
}
// acc + w_i w_j (f(x+,y+) + f(x-,y+) + f(x+,y-) + f(x-,y-)) for the record at shared address `rec`, the four
// reflections evaluated side by side
template <class T, class F, bool CMPL> __device__ __forceinline__ T warp_cell(unsigned rec, const T* p, const Col2<T>& c, T acc)
{
    T bx, ax, xp, xm, v[4];
    lds_rec(rec, bx, ax, xp, xm);  // what the body does not read is dropped by ptxas
    const T xs[4] = {xp, xm, xp, xm}, ys[4] = {c.yp, c.yp, c.ym, c.ym};
    if constexpr (CMPL) {
        const T b_[4] = {bx, ax, bx, ax}, a_[4] = {ax, bx, ax, bx}, d_[4] = {c.dyc, c.dyc, c.cy, c.cy}, c_[4] = {c.cy, c.cy, c.dyc, c.dyc};
        eval2dc_n<T, F, true, 4>(xs, ys, b_, a_, d_, c_, p, v);
    } else {
        eval2d_n<T, F, true, 4, UsesExpTab<T, F>::value ? EXP_SMEM : 0>(xs, ys, p, v);
    }
    return fm::fma(lds_rec_w(rec, T()) * c.wj, (v[0] + v[1]) + (v[2] + v[3]), acc);
}
// Levels 0 .. 3 in ONE pass.  The level-3 slice (16 nodes per half-axis) holds the nodes of levels 0, 1, 2 at its
// indices 7/15, 3/11, 1/5/9/13 (node t of level l is node 2t of level l + 1: t·h_l and 2t·h_(l+1) are the same
// number, caches.jl:170-176).  At these sizes a level costs a warp far more in staging, reduction and stop test
// (~430 instructions) than in cells (12 / 48 / 192 new cells over 32 lanes), so the warp evaluates the whole
// 33 x 33 grid once, keeps the totals of the four levels apart, and then runs the stop tests of adaptive.jl:211-219
// on them: value, error estimate and stop level are those of the level-by-level loop (an integral that stops at
// level 1 or 2 has paid for cells it did not need — still fewer instructions than two or three separate levels).
//   lane = 16 g + j owns column j and rows 8g .. 8g+7, walked in LEVEL order (7 | 3 | 1, 5 | 0, 2, 4, 6): the
//   running sum after the rows of level m is the lane's share of the cells {level(i) <= m} of its column, which
//   belong to total m iff m >= level(j).  Lanes of g = 0 also take the centre column at row j, lanes of g = 1 the
//   centre row at column j: the axis terms, at the level of node j.
template <class T, class F, bool CMPL> __device__ __forceinline__ void warp_levels_0_to_3(const WarpSm<T>& s, const T* p, int lane, T (&S)[4])
{
    constexpr unsigned RB = (unsigned)(WarpRec<T>::ELEMS * sizeof(T));
    const unsigned x = smem_addr_pinned(s.xrec);
    const int g = lane >> 4, j = lane & 15;
    const int ctr = s.nc - 1;
    const Col2<T> ce = warp_col<T, CMPL>(s, g ? j : ctr);
    T acc = warp_cell<T, F, CMPL>(x + (unsigned)(g ? ctr : j) * RB, p, ce, num<T>::zero());
    const Col2<T> c = warp_col<T, CMPL>(s, j);
    const unsigned base = x + (unsigned)(8 * g) * RB;
    acc = warp_cell<T, F, CMPL>(base + 7u * RB, p, c, acc);
    const T P0 = acc;
    acc = warp_cell<T, F, CMPL>(base + 3u * RB, p, c, acc);
    const T P1 = acc;
#pragma unroll 1
    for (unsigned r = base + RB; r < base + 8u * RB; r += 4u * RB) acc = warp_cell<T, F, CMPL>(r, p, c, acc);
    const T P2 = acc;
#pragma unroll 1
    for (unsigned r = base; r < base + 8u * RB; r += 2u * RB) acc = warp_cell<T, F, CMPL>(r, p, c, acc);
    const int tz = __ffs(j + 1) - 1;      // node j is node (j + 1) h_3
    const int lj = tz >= 3 ? 0 : 3 - tz;  // the level that brought it in
    const T zero = num<T>::zero();
    S[0] = lj <= 0 ? P0 : zero;
    S[1] = lj <= 1 ? P1 : zero;
    S[2] = lj <= 2 ? P2 : zero;
    S[3] = acc;
}
// New cells of a level with n >= 32 nodes (0-based: rows i even -> every column, rows i odd -> even columns) and
// its axis terms (centre column x new rows on lanes 0-15, centre row x new columns on lanes 16-31).  N > 0: n at
// compile time, the column / row-group arithmetic folds into constants.
template <class T, class F, bool CMPL, int N> __device__ __forceinline__ T warp_level_sum(const WarpSm<T>& s, const T* p, int n_runtime, int lane)
{
    constexpr unsigned RB = (unsigned)(WarpRec<T>::ELEMS * sizeof(T));
    const int n = N > 0 ? N : n_runtime;
    const unsigned x = smem_addr_pinned(s.xrec), xend = x + (unsigned)n * RB;
    T acc = num<T>::zero();
    for (int j = lane; j < n; j += 32) {  // even rows x all columns
        const Col2<T> c = warp_col<T, CMPL>(s, j);
#pragma unroll 1
        for (unsigned r = x; r < xend; r += 2u * RB) acc = warp_cell<T, F, CMPL>(r, p, c, acc);
    }
    {   // odd rows x even columns: n/2 columns, two row groups when they are 16
        const int half = n >> 1, cpl = half < 32 ? half : 32, sh = __ffs(cpl) - 1, groups = 32 >> sh;
        const unsigned r0 = x + (unsigned)(1 + 2 * (lane >> sh)) * RB, step = (unsigned)(2 * groups) * RB;
        for (int j = (lane & (cpl - 1)) << 1; j < n; j += 64) {
            const Col2<T> c = warp_col<T, CMPL>(s, j);
#pragma unroll 1
            for (unsigned r = r0; r < xend; r += step) acc = warp_cell<T, F, CMPL>(r, p, c, acc);
        }
    }
    const int g = lane >> 4;
    for (int q = lane & 15; q < (n >> 1); q += 16) {  // axis terms of the new nodes 2q
        const Col2<T> ce = warp_col<T, CMPL>(s, g ? 2 * q : s.nc - 1);
        acc = warp_cell<T, F, CMPL>(x + (unsigned)(g ? s.nc - 1 : 2 * q) * RB, p, ce, acc);
    }
    return acc;
}

// adaptive_integrate_2D_avx (+ tensor-complement extension) by one warp, levels 0 .. level_cap with
// 4 <= level_cap, level_cap <= a.max_levels (k_adaptive2d_warp sends shallower calls through adaptive2d_group)
template <class T, class F> __device__ __forceinline__ void adaptive2d_warp(const Args<T>& a, long long b, int lane, unsigned char* mine, int ncap, int level_cap)
{
    constexpr bool CMPL = is_cmpl2d<F>::value;
    constexpr int RE = WarpRec<T>::ELEMS;
    WarpSm<T> s;
    s.nc = ncap + 1;
    s.xrec = reinterpret_cast<T*>(mine);
    s.ycol = s.xrec + (size_t)s.nc * RE;
    unsigned char* raw0 = mine + (size_t)s.nc * RE * sizeof(T) + ((5 * (size_t)s.nc * sizeof(T) + 15) & ~(size_t)15);
    const size_t raw_bytes = (size_t)ncap * node_bytes<T, CMPL>();
    // two mbarriers behind the raw slices; level l uses buffer and barrier l & 1, for the ((l - 3) >> 1)-th time
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(raw0 + 2 * raw_bytes), rawa = (unsigned)__cvta_generic_to_shared(raw0);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // node slice of `level` (n = 2^(level+1) records at offset n - 4) -> raw buffer level & 1, by the TMA unit
    auto prefetch = [&](int level) {
        if (lane != 0) return;
        const int n = 2 << level;
        const void* src = CMPL ? (const void*)(a.nodes_c + (n - 4)) : (const void*)(a.nodes + (n - 4));
        const unsigned bytes = (unsigned)n * node_bytes<T, CMPL>(), bar = bar0 + 8u * (unsigned)(level & 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rawa + (unsigned)(level & 1) * (unsigned)raw_bytes),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
    };
    auto wait_slice = [&](int level) {
        const unsigned bar = bar0 + 8u * (unsigned)(level & 1), parity = (unsigned)((level - 3) >> 1) & 1u;
        unsigned done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        } while (!done);
    };
    prefetch(3);  // both slices are requested before the box is: one global-memory latency instead of two
    prefetch(4);
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    T lo[2], hi[2];
    bool neg;
    int flags;
    if (!load_box<2>(a, b, lo, hi, neg, flags)) {
        if (lane == 0) store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
        wait_slice(3);  // the requested slices land in this warp's shared memory: not before that may it be given up
        wait_slice(4);
        return;
    }
    const T half = num<T>::half();
    s.dx = half * (hi[0] - lo[0]); s.x0 = half * (hi[0] + lo[0]);
    s.dy = half * (hi[1] - lo[1]); s.y0 = half * (hi[1] + lo[1]);
    s.w0 = num<T>::half_pi();
    T h = a.tm * half;
    // levels 0 .. 3 on the level-3 grid, their four stop tests afterwards
    wait_slice(3);
    warp_stage<T, CMPL>(s, raw0 + raw_bytes, 16, lane);
    if (lane == 16) warp_stage_centre<T, CMPL>(s);
    __syncwarp();
    T old_res, err = num<T>::zero();
    Comp<T> s_total;
    {
        T S[4];
        warp_levels_0_to_3<T, F, CMPL>(s, p, lane, S);
        T fc;
        if constexpr (CMPL) fc = (s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, s.dx, s.dx, s.dy, s.dy, p);
        else fc = (s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, p);
        const int m = lane >> 3;  // this lane's class holds the total of level m
        const T tot = warp_reduce4<T>(S, lane) + fc;
        T hm = h;
        for (int q = 0; q < m; ++q) hm = hm * half;
        const T res = s.dx * s.dy * (hm * hm) * tot;
        const T prev = __shfl_up_sync(0xffffffffu, res, 8);
        const T errm = fm::abs(res - prev);
        const bool conv = m >= 1 && errm <= fm::err_target(res, a.rtol, a.atol) && may_stop(a.options, m, a.max_levels);
        const unsigned stop = __ballot_sync(0xffffffffu, conv);
        if (stop) {  // the first level whose test holds is where the level-by-level loop returns
            wait_slice(4);  // the slice of level 4 has landed
            if (lane == __ffs(stop) - 1) store_result(a, b, neg ? -res : res, errm, m, flags | FLAG_CONVERGED);
            return;
        }
        s_total = comp_of(__shfl_sync(0xffffffffu, tot, 24));
        old_res = __shfl_sync(0xffffffffu, res, 24);
        h = __shfl_sync(0xffffffffu, hm, 24);
    }
    int level_used = 3;
    int in_flight = 4;  // level whose slice has been requested but not yet consumed (0: none)
    bool queued = false;
    for (int level = 4; level <= a.max_levels; ++level) {
        if (level > level_cap) {  // too deep for this warp's shared-memory slice: whole-CTA pass
            if (lane == 0) a.tail_idx[atomicAdd(a.tail_count, 1)] = (int)b;
            queued = true;
            break;
        }
        h = h * half;
        const int n = 2 << level;
        __syncwarp();               // every lane is done with the previous level's records
        wait_slice(level);          // the slice of this level has landed in shared memory
        in_flight = 0;
        warp_stage<T, CMPL>(s, raw0 + (size_t)(level & 1) * raw_bytes, n, lane);
        __syncwarp();
        if (level + 1 <= level_cap) {  // the other raw buffer was last read two levels ago: refill it during this level's evaluation
            prefetch(level + 1);
            in_flight = level + 1;
        }
        const T acc = n == 32 ? warp_level_sum<T, F, CMPL, 32>(s, p, 32, lane) : warp_level_sum<T, F, CMPL, 0>(s, p, n, lane);
        const Comp<T> lv_sum = group_reduce<T, 32>(comp_of(acc), nullptr);
        int done = 0;
        if (lane == 0) {
            s_total.merge(lv_sum);
            const T new_res = s.dx * s.dy * (h * h) * s_total.value();
            err = fm::abs(new_res - old_res);
            old_res = new_res;
            level_used = level;
            const bool conv = err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels);
            if (conv) flags |= FLAG_CONVERGED;
            done = conv ? 1 : 0;
        }
        if (__shfl_sync(0xffffffffu, done, 0)) break;
    }
    if (in_flight) wait_slice(in_flight);  // a slice requested for a level that is never evaluated is still consumed
    if (lane == 0 && !queued) {
        if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
}

// Launch shapes of the cooperative 2D adaptive path:
//   k_adaptive2d_warp  one WARP per integral up to level a.level (shallow integrals: 4 225 evaluations at
//                      level 4 keep 32 lanes busy without any CTA barrier); CTAs of 4 warps and at most 72
//                      registers, so that 7 CTAs = 28 warps fit an SM and 4096 integrals run as ONE wave;
//                      integrals that need deeper levels are queued (a.tail_idx)
//   k_adaptive2d_cta   one CTA per integral: blockIdx = integral, or — with a queue in a.tail_idx —
//                      the CTAs walk the queued integrals (second launch after k_adaptive2d_warp)
constexpr int WARP2D_WARPS = 4;
template <class T, class F> __global__ void __launch_bounds__(WARP2D_WARPS * 32, 7) k_adaptive2d_warp(const Args<T> a)
{
    pdl_launch_dependents();     // the queue pass (k_adaptive2d_cta) may be set up behind this grid
    family_tables_init<T, F>();  // barrier inside: before any early return (constant tables only)
    constexpr bool CMPL = is_cmpl2d<F>::value;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * WARP2D_WARPS + warp;
    const int ncap = 2 << a.level;
    unsigned char* mine = fts_dyn_smem + (size_t)warp * warp2d_bytes<T, CMPL>(ncap);
    // This grid is itself launched as a programmatic dependent of whatever precedes it on the stream (in a loop of
    // calls: the previous call's queue pass): its CTAs are set up behind that grid and touch global memory only
    // from here on, after its completion and memory flush.
    pdl_wait_primary();
    if (b >= a.batch) return;
    if (a.level >= 4) {
        adaptive2d_warp<T, F>(a, b, lane, mine, ncap, a.level);
        return;
    }
    // max_levels < 4 (or a lower cap asked for): the level-by-level group loop
    unsigned char* raw0 = mine + sm2_elems<T, CMPL>(ncap) * sizeof(T);
    NodePipe np;
    pipe_init(np, raw0 + 2 * (size_t)ncap * node_bytes<T, CMPL>(), raw0, raw0 + (size_t)ncap * node_bytes<T, CMPL>(), lane == 0);
    __syncwarp();
    adaptive2d_group<T, F, 32>(a, b, lane, reinterpret_cast<T*>(mine), ncap, a.level, nullptr, nullptr, np);
}

template <class T, class F> __global__ void __launch_bounds__(CTA_THREADS) k_adaptive2d_cta(const Args<T> a)
{
    constexpr bool CMPL = is_cmpl2d<F>::value;
    __shared__ T red[64];
    __shared__ int s_done;
    if (a.tail_idx != nullptr) {
        // queue pass behind the warp pass: most calls leave nothing in the queue, and an empty pass should cost a
        // launch, not a table set-up (the count is uniform over the grid)
        pdl_launch_dependents();  // the next call's warp pass may be set up behind this grid
        pdl_wait_primary();  // the warp pass has completed: the queue is final and visible
        if (*a.tail_count == 0) return;
    }
    family_tables_init<T, F>();  // barrier inside: before any divergent return
    unsigned char* dyn = fts_dyn_smem;
    NodePipe np;
    const int nmax = a.max_levels > 0 ? (2 << a.max_levels) : 2;
    T* smem = reinterpret_cast<T*>(dyn);
    unsigned char* raw0 = dyn + sm2_elems<T, CMPL>(nmax) * sizeof(T);
    pipe_init(np, raw0 + 2 * (size_t)nmax * node_bytes<T, CMPL>(), raw0, raw0 + (size_t)nmax * node_bytes<T, CMPL>(), threadIdx.x == 0);
    __syncthreads();
    if (a.tail_idx == nullptr) {
        adaptive2d_group<T, F, CTA_THREADS>(a, blockIdx.x, threadIdx.x, smem, nmax, a.max_levels, red, &s_done, np);
        return;
    }
    const int count = *a.tail_count;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        adaptive2d_group<T, F, CTA_THREADS>(a, a.tail_idx[q], threadIdx.x, smem, nmax, a.max_levels, red, &s_done, np);
        __syncthreads();
    }
    if (threadIdx.x == 0) {  // the last CTA to leave re-arms the queue (see k_adaptive1d_tail)
        __threadfence();
        const unsigned int t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {
            *a.tail_count = 0;
            *a.ticket = 0u;
        }
    }
}

// integrate2D_avx   integrate2D.jl:46-73
template <class T, class F> __global__ void __launch_bounds__(CTA_THREADS) k_fixed2d_cta(const Args<T> a)
{
    family_tables_init<T, F>();  // barrier inside: before any early return
    __shared__ T red[64];
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T* lo = a.low + b * a.bstride;
    const T* hi = a.up + b * a.bstride;
    Sm2<T> s;
    s.dx = half * (hi[0] - lo[0]); s.x0 = half * (hi[0] + lo[0]);
    s.dy = half * (hi[1] - lo[1]); s.y0 = half * (hi[1] + lo[1]);
    s.w0 = num<T>::half_pi();
    sm2_carve<T, false>(s, reinterpret_cast<T*>(fts_dyn_smem), a.n);
    sm2_stage<T, false>(s, &a.grid->x, &a.grid->w, nullptr, 2, a.n);
    __syncthreads();
    T acc = level_sum_2d<T, F, false>(s, p, a.n, 0, true, tid, blockDim.x);
    if (tid == 0) acc = acc + (s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, p);
    const Comp<T> tot = block_reduce(comp_of(acc), red);
    if (tid == 0) a.value[b] = s.dx * s.dy * (a.h * a.h) * tot.value();
}

// ------------------------------------------------------------------------------------ 3D -----
template <class T> struct Sm3 {
    T *xp, *xm, *yp, *ym, *zp, *zm, *w;
    // z coordinates and weights once more, permuted k -> (k mod 4)(n/4) + k/4: the lanes of a warp
    // walk runs of 4 consecutive k (cells3_run), lane m at k = 4m + c, and read element c(n/4) + m —
    // consecutive words, no bank conflicts (stride-4 reads of the natural layout are 4-way)
    T *zpq, *zmq, *wq;
    T dx, x0, dy, y0, dz, z0, w0;
};
constexpr int SM3_ARRAYS = 10;
__device__ __forceinline__ int sm3_q(int k, int n) { return (k & 3) * (n >> 2) + (k >> 2); }

template <class T> __device__ __forceinline__ void sm3_carve(Sm3<T>& s, T* base, int n)
{
    s.xp = base; s.xm = base + n; s.yp = base + 2 * n; s.ym = base + 3 * n; s.zp = base + 4 * n; s.zm = base + 5 * n; s.w = base + 6 * n;
    s.zpq = base + 7 * n; s.zmq = base + 8 * n; s.wq = base + 9 * n;
}

// the reference's coordinate hoist (adaptive.jl:570-580), into shared memory
template <class T> __device__ __forceinline__ void sm3_stage(Sm3<T>& s, const T* x, const T* w, int stride, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const T xi = x[(size_t)i * stride];
        s.xp[i] = fm::fma(s.dx, xi, s.x0);
        s.xm[i] = fm::fma(-s.dx, xi, s.x0);
        s.yp[i] = fm::fma(s.dy, xi, s.y0);
        s.ym[i] = fm::fma(-s.dy, xi, s.y0);
        s.zp[i] = fm::fma(s.dz, xi, s.z0);
        s.zm[i] = fm::fma(-s.dz, xi, s.z0);
        s.w[i] = w[(size_t)i * stride];
        const int q = sm3_q(i, n);
        s.zpq[q] = s.zp[i];
        s.zmq[q] = s.zm[i];
        s.wq[q] = s.w[i];
    }
}

// Families with an interleaved batch evaluator (exp) get their 8 / 12 / 6 reflections evaluated
// side by side; every other functor keeps the direct expression (the compiler then shares
// sub-expressions between reflections, e.g. 6 logs instead of 24 for log(1-x)log(1-y)log(1-z)).
#define FTS_FE(X, Y, Z) F::template eval<T, true>(X, Y, Z, p)
template <class T, class F> __device__ __forceinline__ T cell3(const Sm3<T>& s, const T* p, int i, int j, int k)
{
    const T xp = s.xp[i], xm = s.xm[i], yp = s.yp[j], ym = s.ym[j], zp = s.zp[k], zm = s.zm[k];
    T f;
    if constexpr (HasEvalN<F>::value) {
        const T xs[8] = {xp, xm, xp, xm, xp, xm, xp, xm}, ys[8] = {yp, yp, ym, ym, yp, yp, ym, ym}, zs[8] = {zp, zp, zp, zp, zm, zm, zm, zm};
        T v[8];
        eval3d_n<T, F, true, 8, UsesExpTab<T, F>::value ? EXP_SMEM : 0>(xs, ys, zs, p, v);
        f = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    } else {
        f = ((FTS_FE(xp, yp, zp) + FTS_FE(xm, yp, zp)) + (FTS_FE(xp, ym, zp) + FTS_FE(xm, ym, zp))) +
            ((FTS_FE(xp, yp, zm) + FTS_FE(xm, yp, zm)) + (FTS_FE(xp, ym, zm) + FTS_FE(xm, ym, zm)));
    }
    return (s.w[i] * s.w[j] * s.w[k]) * f;
}
// CH cells along k (k0, k0+kstep, ...) that share (i, j): the x/y coordinates, w_i*w_j and whatever
// the integrand computes from x and y alone (log(1-x)log(1-y)...) are formed once per run instead
// of once per cell, and the flat-index decode is paid once per CH cells.
template <class T, class F, int CH> __device__ __forceinline__ T cells3_run(const Sm3<T>& s, const T* p, int n, int i, int j, int k0, int kstep)
{
    const T xp = s.xp[i], xm = s.xm[i], yp = s.yp[j], ym = s.ym[j];
    if constexpr (HasCellsRun<F>::value) {
        T zp[CH], zm[CH], wk[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int q = sm3_q(k0 + c * kstep, n);
            zp[c] = s.zpq[q];
            zm[c] = s.zmq[q];
            wk[c] = s.wq[q];
        }
        return (s.w[i] * s.w[j]) * F::template cells_run<T, CH>(xp, xm, yp, ym, zp, zm, wk, p);
    }
    T acc = num<T>::zero();
    // interleaved-batch integrands already keep 8 chains in flight per cell: unrolling the run as
    // well costs 30+ registers (156 -> 1 CTA/SM); everyone else is unrolled so that common
    // sub-expressions of the CH cells merge
    constexpr int UNROLL = HasEvalN<F>::value ? 1 : CH;
#pragma unroll UNROLL
    for (int c = 0; c < CH; ++c) {
        const int q = sm3_q(k0 + c * kstep, n);
        const T zp = s.zpq[q], zm = s.zmq[q];
        T f;
        if constexpr (HasEvalN<F>::value) {
            const T xs[8] = {xp, xm, xp, xm, xp, xm, xp, xm}, ys[8] = {yp, yp, ym, ym, yp, yp, ym, ym}, zs[8] = {zp, zp, zp, zp, zm, zm, zm, zm};
            T v[8];
            eval3d_n<T, F, true, 8, UsesExpTab<T, F>::value ? EXP_SMEM : 0>(xs, ys, zs, p, v);
            f = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        } else {
            f = ((FTS_FE(xp, yp, zp) + FTS_FE(xm, yp, zp)) + (FTS_FE(xp, ym, zp) + FTS_FE(xm, ym, zp))) +
                ((FTS_FE(xp, yp, zm) + FTS_FE(xm, yp, zm)) + (FTS_FE(xp, ym, zm) + FTS_FE(xm, ym, zm)));
        }
        acc = fm::fma(s.wq[q], f, acc);
    }
    return (s.w[i] * s.w[j]) * acc;
}
template <class T, class F> __device__ __forceinline__ T plane3(const Sm3<T>& s, const T* p, int i, int j)
{
    const T x0 = s.x0, y0 = s.y0, z0 = s.z0;
    const T xp = s.xp[i], xm = s.xm[i], ypi = s.yp[i], ymi = s.ym[i];
    const T ypj = s.yp[j], ymj = s.ym[j], zpj = s.zp[j], zmj = s.zm[j];
    T f;
    if constexpr (HasEvalN<F>::value) {
        const T xs[12] = {xp, xm, xp, xm, xp, xm, xp, xm, x0, x0, x0, x0};
        const T ys[12] = {ypj, ypj, ymj, ymj, y0, y0, y0, y0, ypi, ymi, ypi, ymi};
        const T zs[12] = {z0, z0, z0, z0, zpj, zpj, zmj, zmj, zpj, zpj, zmj, zmj};
        T v[12];
        eval3d_n<T, F, true, 12, UsesExpTab<T, F>::value ? EXP_SMEM : 0>(xs, ys, zs, p, v);
        f = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7])) + ((v[8] + v[9]) + (v[10] + v[11]));
    } else {
        f = ((FTS_FE(xp, ypj, z0) + FTS_FE(xm, ypj, z0)) + (FTS_FE(xp, ymj, z0) + FTS_FE(xm, ymj, z0))) +
            ((FTS_FE(xp, y0, zpj) + FTS_FE(xm, y0, zpj)) + (FTS_FE(xp, y0, zmj) + FTS_FE(xm, y0, zmj))) +
            ((FTS_FE(x0, ypi, zpj) + FTS_FE(x0, ymi, zpj)) + (FTS_FE(x0, ypi, zmj) + FTS_FE(x0, ymi, zmj)));
    }
    return (s.w[i] * s.w[j] * s.w0) * f;
}
template <class T, class F> __device__ __forceinline__ T axis3(const Sm3<T>& s, const T* p, int i)
{
    const T x0 = s.x0, y0 = s.y0, z0 = s.z0;
    const T f = (FTS_FE(s.xp[i], y0, z0) + FTS_FE(s.xm[i], y0, z0)) + (FTS_FE(x0, s.yp[i], z0) + FTS_FE(x0, s.ym[i], z0)) +
                (FTS_FE(x0, y0, s.zp[i]) + FTS_FE(x0, y0, s.zm[i]));
    return (s.w[i] * (s.w0 * s.w0)) * f;
}
#undef FTS_FE

// New cells of a level in 0-based indices (adaptive.jl:430-456 / 585-690):
//   class A  i even           : all (j,k) octants, planes(i,j) for all j, axes(i)
//   class B  i odd, j even    : all k octants, planes(i,j)
//   class C  i odd, j odd     : k even octants only
template <class T, class F>
__device__ __forceinline__ T level_sum_3d(const Sm3<T>& s, const T* p, int n, int log2n, bool all_new, long long start, long long stride)
{
    T acc = num<T>::zero();
    if (all_new) {
        const long long n2 = (long long)n * n, ncell = n2 * n;
        for (long long t = start; t < ncell; t += stride) {
            const int i = (int)(t / n2);
            const int r = (int)(t - (long long)i * n2);
            const int j = r / n, k = r - j * n;
            acc = acc + cell3<T, F>(s, p, i, j, k);
        }
        for (long long t = start; t < n2; t += stride) {
            const int i = (int)(t / n), j = (int)(t - (long long)i * n);
            acc = acc + plane3<T, F>(s, p, i, j);
        }
        for (long long t = start; t < n; t += stride) acc = acc + axis3<T, F>(s, p, (int)t);
    } else {
        const int l2 = log2n;
        const long long n3 = 1LL << (3 * l2);
        const long long nA = n3 >> 1, nB = n3 >> 2, nC = n3 >> 3;
        const long long ncell = nA + nB + nC;
        // Runs of CH cells along k when every thread gets several of them (rows hold n, or n/2 in
        // class C, cells and every class size is a multiple of CH for n >= 2 CH).
        constexpr int CH = 4;
        const bool runs = n >= 2 * CH && ncell / CH >= 4 * stride;
        for (long long u4 = start; runs && u4 < ncell / CH; u4 += stride) {
            const long long t = u4 * CH;
            int i, j, k, kstep = 1;
            if (t < nA) {
                i = (int)(t >> (2 * l2)) << 1;
                j = (int)(t >> l2) & (n - 1);
                k = (int)t & (n - 1);
            } else if (t < nA + nB) {
                const long long u = t - nA;
                i = ((int)(u >> (2 * l2 - 1)) << 1) | 1;
                const int r = (int)(u & ((1LL << (2 * l2 - 1)) - 1));
                j = (r >> l2) << 1;
                k = r & (n - 1);
            } else {
                const long long u = t - nA - nB;
                i = ((int)(u >> (2 * l2 - 2)) << 1) | 1;
                const int r = (int)(u & ((1LL << (2 * l2 - 2)) - 1));
                j = ((r >> (l2 - 1)) << 1) | 1;
                k = (r & ((n >> 1) - 1)) << 1;
                kstep = 2;
            }
            acc = acc + cells3_run<T, F, CH>(s, p, n, i, j, k, kstep);
        }
        for (long long t = start; !runs && t < ncell; t += stride) {
            int i, j, k;
            if (t < nA) {
                i = (int)(t >> (2 * l2)) << 1;
                j = (int)(t >> l2) & (n - 1);
                k = (int)t & (n - 1);
            } else if (t < nA + nB) {
                const long long u = t - nA;
                i = ((int)(u >> (2 * l2 - 1)) << 1) | 1;
                const int r = (int)(u & ((1LL << (2 * l2 - 1)) - 1));
                j = (r >> l2) << 1;
                k = r & (n - 1);
            } else {
                const long long u = t - nA - nB;
                i = ((int)(u >> (2 * l2 - 2)) << 1) | 1;
                const int r = (int)(u & ((1LL << (2 * l2 - 2)) - 1));
                j = ((r >> (l2 - 1)) << 1) | 1;
                k = (r & ((n >> 1) - 1)) << 1;
            }
            acc = acc + cell3<T, F>(s, p, i, j, k);
        }
        const long long pA = ((long long)n * n) >> 1, nplane = pA + (pA >> 1);
        for (long long t = start; t < nplane; t += stride) {
            int i, j;
            if (t < pA) {
                i = (int)(t >> l2) << 1;
                j = (int)(t & (n - 1));
            } else {
                const long long u = t - pA;
                i = ((int)(u >> (l2 - 1)) << 1) | 1;
                j = (int)(u & ((n >> 1) - 1)) << 1;
            }
            acc = acc + plane3<T, F>(s, p, i, j);
        }
        for (long long t = start; t < (n >> 1); t += stride) acc = acc + axis3<T, F>(s, p, (int)t << 1);
    }
    return acc;
}

// adaptive_integrate_3D_avx   adaptive.jl:489-706
template <class T, class F> __global__ void __launch_bounds__(CTA_THREADS) k_adaptive3d_cta(const Args<T> a)
{
    family_tables_init<T, F>();  // barrier inside: before any early return
    __shared__ T red[64];
    __shared__ int s_done;
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    T lo[3], hi[3];
    bool neg;
    int flags;
    if (!load_box<3>(a, b, lo, hi, neg, flags)) {
        if (tid == 0) store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
        return;
    }
    const T half = num<T>::half();
    Sm3<T> s;
    s.dx = half * (hi[0] - lo[0]); s.x0 = half * (hi[0] + lo[0]);
    s.dy = half * (hi[1] - lo[1]); s.y0 = half * (hi[1] + lo[1]);
    s.dz = half * (hi[2] - lo[2]); s.z0 = half * (hi[2] + lo[2]);
    s.w0 = num<T>::half_pi();
    const int nmax = a.max_levels > 0 ? (2 << a.max_levels) : 2;
    sm3_carve<T>(s, reinterpret_cast<T*>(fts_dyn_smem), nmax);
    T h = a.tm * half;
    sm3_stage<T>(s, a.ix, a.iw, 1, 2);
    __syncthreads();
    T acc = level_sum_3d<T, F>(s, p, 2, 1, true, tid, blockDim.x);
    if (tid == 0) acc = acc + (s.w0 * s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, s.z0, p);
    Comp<T> s_total = block_reduce(comp_of(acc), red);
    T old_res = s.dx * s.dy * s.dz * (h * h * h) * s_total.value();
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        const int n = 2 << level;
        const Node<T>* lv = a.nodes + (n - 4);
        sm3_stage<T>(s, &lv->x, &lv->w, 2, n);
        __syncthreads();
        acc = level_sum_3d<T, F>(s, p, n, level + 1, false, tid, blockDim.x);
        const Comp<T> lv_sum = block_reduce(comp_of(acc), red);
        if (tid == 0) {
            s_total.merge(lv_sum);
            const T new_res = s.dx * s.dy * s.dz * (h * h * h) * s_total.value();
            err = fm::abs(new_res - old_res);
            old_res = new_res;
            level_used = level;
            const bool conv = err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels);
            if (conv) flags |= FLAG_CONVERGED;
            s_done = conv ? 1 : 0;
        }
        __syncthreads();
        if (s_done) break;
    }
    if (tid == 0) {
        if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
}

// integrate3D_avx   integrate3D.jl:68-123
template <class T, class F> __global__ void __launch_bounds__(CTA_THREADS) k_fixed3d_cta(const Args<T> a)
{
    family_tables_init<T, F>();  // barrier inside: before any early return
    __shared__ T red[64];
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T* lo = a.low + b * a.bstride;
    const T* hi = a.up + b * a.bstride;
    Sm3<T> s;
    s.dx = half * (hi[0] - lo[0]); s.x0 = half * (hi[0] + lo[0]);
    s.dy = half * (hi[1] - lo[1]); s.y0 = half * (hi[1] + lo[1]);
    s.dz = half * (hi[2] - lo[2]); s.z0 = half * (hi[2] + lo[2]);
    s.w0 = num<T>::half_pi();
    sm3_carve<T>(s, reinterpret_cast<T*>(fts_dyn_smem), a.n);
    sm3_stage<T>(s, &a.grid->x, &a.grid->w, 2, a.n);
    __syncthreads();
    T acc = level_sum_3d<T, F>(s, p, a.n, 0, true, tid, blockDim.x);
    if (tid == 0) acc = acc + (s.w0 * s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, s.z0, p);
    const Comp<T> tot = block_reduce(comp_of(acc), red);
    if (tid == 0) a.value[b] = (s.dx * s.dy * s.dz * (a.h * a.h * a.h)) * tot.value();
}

// ---------------------------------------------------------------- generic n-D tensor product ----
// integrate1D/2D/3D's weighted node sum on a D-dimensional box (the reference's stated next step, JOSS paper.md:121):
// sum over the (2n+1)^D points of prod_d w_{i_d} f(x0 + dx * xi_{i_d}), times prod_d dx_d * h^D.  The signed node
// table (index 0 = centre, 2k-1 / 2k = +-x_k) sits in shared memory; gridDim.x CTAs per integral (blockIdx.y) walk
// the flat point index with a stride, carrying the D base-(2n+1) digits instead of dividing; compensated partials
// per CTA, the last CTA of an integral merges them in CTA order.  F::eval(const T* x, const T* p).
template <class T, class F, int D> __global__ void __launch_bounds__(CTA_THREADS) k_fixed_nd(const Args<T> a)
{
    __shared__ T red[64];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const long long b = blockIdx.y;
    const int m = 2 * a.n + 1;
    T* xs = reinterpret_cast<T*>(fts_dyn_smem);
    T* ws = xs + m;
    for (int i = tid; i < m; i += blockDim.x) {
        if (i == 0) {
            xs[0] = num<T>::zero();
            ws[0] = num<T>::half_pi();
        } else {
            const Node<T> nd = a.grid[(i - 1) >> 1];
            xs[i] = (i & 1) ? nd.x : -nd.x;
            ws[i] = nd.w;
        }
    }
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    T dx[D], x0[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const T lo = a.low[b * a.bstride + d], hi = a.up[b * a.bstride + d];
        dx[d] = half * (hi - lo);
        x0[d] = half * (hi + lo);
    }
    __syncthreads();
    // work unit = (outer point: axes 1 ... D-1, chunk of <= 32 points along axis 0): the outer coordinates and
    // weight product are formed once per unit, the flat unit index is walked by carrying its digits
    // (digit 0 = chunk, base nch; digits 1 ... D-1 = outer axes, base m) instead of dividing
    constexpr int ND_RUN_MAX = 32;
    const int nch = (m + ND_RUN_MAX - 1) / ND_RUN_MAX;
    const int ND_RUN = (m + nch - 1) / nch;  // equal chunks (161 points: 6 x 27, not 5 x 32 + 1)
    unsigned long long total = (unsigned long long)nch;
#pragma unroll
    for (int d = 1; d < D; ++d) total *= (unsigned long long)m;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + tid;
    int dig[D], sdig[D];
    {
        unsigned long long q = t, r = stride;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const unsigned long long base = d == 0 ? (unsigned long long)nch : (unsigned long long)m;
            dig[d] = (int)(q % base); q /= base;
            sdig[d] = (int)(r % base); r /= base;
        }
    }
    Comp<T> acc;
    for (; t < total; t += stride) {
        T x[D];
        T wo = num<T>::one();
#pragma unroll
        for (int d = 1; d < D; ++d) {
            x[d] = fm::fma(dx[d], xs[dig[d]], x0[d]);
            wo = wo * ws[dig[d]];
        }
        const int i0 = dig[0] * ND_RUN, i1 = i0 + ND_RUN < m ? i0 + ND_RUN : m;
        T run = num<T>::zero();
        for (int i = i0; i < i1; ++i) {
            x[0] = fm::fma(dx[0], xs[i], x0[0]);
            run = fm::fma(ws[i], F::template eval<T, true>(x, p), run);
        }
        acc.add(wo * run);
        int carry = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int base = d == 0 ? nch : m;
            int v = dig[d] + sdig[d] + carry;
            carry = v >= base ? 1 : 0;
            dig[d] = v - (carry ? base : 0);
        }
    }
    const Comp<T> bsum = block_reduce(acc, red);
    T* partials = a.block_partials + 2 * (size_t)b * gridDim.x;
    if (tid == 0) {
        partials[2 * blockIdx.x] = bsum.s;
        partials[2 * blockIdx.x + 1] = bsum.c;
        __threadfence();
        const unsigned int k = atomicAdd(a.ticket + b, 1u);
        s_last = (k == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    Comp<T> tot;
    for (int i = tid; i < (int)gridDim.x; i += blockDim.x) {
        Comp<T> o;
        o.s = __ldcg(partials + 2 * i);
        o.c = __ldcg(partials + 2 * i + 1);
        tot.merge(o);
    }
    tot = block_reduce(tot, red);
    if (tid == 0) {
        T scale = num<T>::one();
#pragma unroll
        for (int d = 0; d < D; ++d) scale = scale * (dx[d] * a.h);
        a.value[b] = scale * tot.value();
        a.ticket[b] = 0u;
    }
}

// Stop test after the partials of all ranks are known (adaptive.jl:693-700): one thread, identical arithmetic
// on every rank.  part(r) = rank r's compensated (s, c) of the level, merged in rank order.
// state = {s_total.s, s_total.c, old_res, value, err, done, level_used, h}
// state[5] (done): 0 running, 1 converged, 2 zero-width box under the quad() edge rules, 3 a peer never delivered.
template <class T, int D, class Part>
__device__ __forceinline__ void level_step_apply(const T* low, const T* up, T tm, T rtol, T atol, int level, int nranks, Part part, T* state, int options, int max_levels)
{
    if (level == 0 && (options & OPT_QUAD_EDGE)) {  // any(low .== up) -> zero(T)   quad.jl:96-98,129-131
        bool zero = false;
#pragma unroll
        for (int d = 0; d < D; ++d) zero = zero || (low[d] == up[d]);
        if (zero) {
#pragma unroll
            for (int i = 0; i < 8; ++i) state[i] = num<T>::zero();
            state[5] = num<T>::from_double(2.0);
            return;
        }
    }
    const T half = num<T>::half();
    Comp<T> tot;
    if (level > 0) {
        tot.s = state[0];
        tot.c = state[1];
    }
    for (int r = 0; r < nranks; ++r) tot.merge(part(r));
    T scale = num<T>::one();
#pragma unroll
    for (int d = 0; d < D; ++d) scale = scale * (half * (up[d] - low[d]));
    const T h = level == 0 ? tm * half : state[7] * half;
    T hD = h;
#pragma unroll
    for (int d = 1; d < D; ++d) hD = hD * h;
    const T res = scale * hD * tot.value();
    state[0] = tot.s;
    state[1] = tot.c;
    state[7] = h;
    state[6] = num<T>::from_double((double)level);
    if (level == 0) {
        state[2] = res;
        state[3] = res;
        state[4] = num<T>::zero();
        state[5] = num<T>::zero();
    } else {
        const T err = fm::abs(res - state[2]);
        state[2] = res;
        state[3] = res;
        state[4] = err;
        if (err <= fm::err_target(res, rtol, atol) && may_stop(options, level, max_levels)) state[5] = num<T>::one();
    }
}

// lanes 0 .. nranks-1 of a warp wait for the peers' flags of `slot_row` in this rank's own mailbox; returns false
// on timeout; *waited_ns = the longest wait
__device__ __forceinline__ bool peer_wait_flags(const double* mail, int slot_row, int nranks, double epoch, unsigned long long timeout_ns, unsigned long long* waited_ns)
{
    const int lane = threadIdx.x & 31;
    bool ok = true;
    unsigned long long waited = 0ULL;
    if (lane < nranks) {
        const volatile double* slot = mail + ((size_t)slot_row * nranks + lane) * 4;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        t1 = t0;
        while (slot[2] != epoch) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                ok = false;
                break;
            }
        }
        waited = t1 - t0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_down_sync(0xffffffffu, waited, off);
        waited = o > waited ? o : waited;
    }
    *waited_ns = __shfl_sync(0xffffffffu, waited, 0);
    ok = __all_sync(0xffffffffu, ok);
    __threadfence_system();
    return ok;
}

// ------------------------------------------------------------------ one level, whole grid ----
// Partial sum of level `a.level` (0 = base grid) of ONE 2D/3D integral (bounds/params at index
// 0), flat new-cell index dealt over gridDim.x CTAs and a.nranks devices.  The last CTA to
// finish (ticket) merges the per-CTA (s,c) partials in CTA order and writes a.partial_out[0..1].
//
// a.step_fused: that last CTA also exchanges the partial with the peers (mailboxes over NVLink), waits for
// theirs and does the stop test of the level, and the kernel walks the levels a.level ... a.level_hi in ONE
// launch.  With more than one CTA (a.gen != nullptr; cooperative launch, every CTA resident) the other CTAs
// wait at a generation barrier for the last CTA's verdict: one launch per integral for the levels whose
// duration is of the order of a launch (levels 0 ... 7 of a 3D integral: 6-10 us each as launches, 2-3 us
// as barrier rounds), which is what bounds the strong scaling of a sharded integral.  Levels below
// a.shard_from are computed whole on every rank (no exchange), the others on this rank's shard.
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
// the generation word holds (count << 1) | done; counts are compared modulo 2^31
__device__ __forceinline__ unsigned int wait_generation(const unsigned int* gen, unsigned int target)
{
    unsigned int v;
    for (;;) {
        v = ld_acquire_gpu(gen);
        if ((int)(((v >> 1) - target) << 1) >= 0) break;
        __nanosleep(40);  // several hundred CTAs poll one word: leave the L2 slice room for the store they wait for
    }
    return v;
}

template <class T, class F, int D> __global__ void __launch_bounds__(CTA_THREADS) k_level_grid(const Args<T> a)
{
    constexpr bool CMPL = is_cmpl2d<F>::value;
    __shared__ T red[64];
    __shared__ int s_last;
    __shared__ unsigned int s_gen;
    if (a.state && a.state[5] != num<T>::zero()) return;  // converged at an earlier level
    family_tables_init<T, F>();
    const int tid = threadIdx.x;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, 0);
    const T half = num<T>::half();
    const int level_hi = a.step_fused ? a.level_hi : a.level;
    for (int level = a.level; level <= level_hi; ++level) {
        if (a.step_fused && level > a.level) {  // converged at an earlier level of this launch (state_out is only written by last CTAs)
            const bool done = a.gen ? (s_gen & 1u) != 0u : __ldcg(a.state_out + 5) != num<T>::zero();
            if (done) {
                if (a.wait_ns && blockIdx.x == 0 && tid == 0) a.wait_ns[level] = 0ULL;
                return;
            }
        }
        const bool sharded = level >= a.shard_from;
        const int rank = sharded ? a.rank : 0, nranks = sharded ? a.nranks : 1;
        // CTAs of the level: the whole grid, or the first range_blocks[level] CTAs of a multi-CTA range (a small level on
        // 592 CTAs would spend its time in 592 tickets to one address and a 592-term merge)
        const unsigned int nblk = a.gen ? min((unsigned int)a.range_blocks[level & 31], gridDim.x) : gridDim.x;
        if (blockIdx.x >= nblk) {
            if (level == level_hi) return;
            if (tid == 0) s_gen = wait_generation(a.gen, a.gen_base + (unsigned int)(level + 1));
            __syncthreads();
            continue;
        }
        const long long gthreads = (long long)nblk * blockDim.x;
        const long long start = (long long)rank * gthreads + (long long)blockIdx.x * blockDim.x + tid;
        const long long stride = gthreads * nranks;
        const int n = level == 0 ? 2 : (2 << level);
        T acc;
        if constexpr (D == 3) {
            Sm3<T> s;
            s.dx = half * (a.up[0] - a.low[0]); s.x0 = half * (a.up[0] + a.low[0]);
            s.dy = half * (a.up[1] - a.low[1]); s.y0 = half * (a.up[1] + a.low[1]);
            s.dz = half * (a.up[2] - a.low[2]); s.z0 = half * (a.up[2] + a.low[2]);
            s.w0 = num<T>::half_pi();
            sm3_carve<T>(s, reinterpret_cast<T*>(fts_dyn_smem), n);
            if (level == 0) sm3_stage<T>(s, a.ix, a.iw, 1, 2);
            else { const Node<T>* lv = a.nodes + (n - 4); sm3_stage<T>(s, &lv->x, &lv->w, 2, n); }
            __syncthreads();
            acc = level_sum_3d<T, F>(s, p, n, level + 1, level == 0, start, stride);
            if (level == 0 && start == 0) acc = acc + (s.w0 * s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, s.z0, p);
        } else {
            Sm2<T> s;
            s.dx = half * (a.up[0] - a.low[0]); s.x0 = half * (a.up[0] + a.low[0]);
            s.dy = half * (a.up[1] - a.low[1]); s.y0 = half * (a.up[1] + a.low[1]);
            s.w0 = num<T>::half_pi();
            sm2_carve<T, CMPL>(s, reinterpret_cast<T*>(fts_dyn_smem), n);
            if (level == 0) sm2_stage<T, CMPL>(s, a.ix, a.iw, a.ic, 1, 2);
            else if constexpr (CMPL) { const NodeC<T>* lv = a.nodes_c + (n - 4); sm2_stage<T, CMPL>(s, &lv->x, &lv->w, &lv->c, 4, n); }
            else { const Node<T>* lv = a.nodes + (n - 4); sm2_stage<T, CMPL>(s, &lv->x, &lv->w, nullptr, 2, n); }
            __syncthreads();
            acc = level_sum_2d<T, F, CMPL>(s, p, n, level + 1, level == 0, start, stride);
            if (level == 0 && start == 0) {
                if constexpr (CMPL) acc = acc + (s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, s.dx, s.dx, s.dy, s.dy, p);
                else {
                    if constexpr (D == 2) acc = acc + (s.w0 * s.w0) * F::template eval<T, true>(s.x0, s.y0, p);
                }
            }
        }
        const Comp<T> bsum = block_reduce(comp_of(acc), red);
        Comp<T> tot;
        if (nblk == 1) {
            // a level that is ONE CTA small (levels 1 .. 4 of a 3D integral): its sum is the level's sum — no partials in
            // global memory, no ticket, no read-back (three dependent global-memory round trips of a ~8 us level)
            tot = bsum;  // valid in thread 0, the only one that uses it
            if (tid == 0) {
                a.partial_out[0] = tot.s;
                a.partial_out[1] = tot.c;
            }
        } else {
            if (tid == 0) {
                a.block_partials[2 * blockIdx.x] = bsum.s;
                a.block_partials[2 * blockIdx.x + 1] = bsum.c;
                const unsigned int t = atom_add_acq_rel_gpu(a.ticket, 1u);  // releases the partials, acquires the other CTAs'
                s_last = (t == nblk - 1) ? 1 : 0;
            }
            __syncthreads();
            if (!s_last) {
                if (!a.step_fused || level == level_hi || a.gen == nullptr) return;
                // generation barrier: the last CTA publishes level + 1 (and the verdict) after the stop test of the level
                if (tid == 0) s_gen = wait_generation(a.gen, a.gen_base + (unsigned int)(level + 1));
                __syncthreads();
                continue;
            }
            for (int i = tid; i < (int)nblk; i += blockDim.x) {
                Comp<T> o;
                o.s = __ldcg(a.block_partials + 2 * i);
                o.c = __ldcg(a.block_partials + 2 * i + 1);
                tot.merge(o);
            }
            tot = block_reduce(tot, red);
            if (tid == 0) {
                a.partial_out[0] = tot.s;
                a.partial_out[1] = tot.c;
                *a.ticket = 0u;
            }
        }
        const bool exchange = a.peer_mail[0] != nullptr && sharded;
        const int slot_row = a.mail_slot + (level - a.level);
        // fused all-gather: one lane per destination rank writes {s, c} into that rank's mailbox over
        // NVLink (or locally for r == rank), fences system-wide, then raises the epoch flag
        if (exchange && tid < nranks) {
            const double s_ = num<T>::to_double(__shfl_sync(0xffffffffu >> (32 - nranks), tot.s, 0));
            const double c_ = num<T>::to_double(__shfl_sync(0xffffffffu >> (32 - nranks), tot.c, 0));
            volatile double* slot = a.peer_mail[tid] + ((size_t)slot_row * nranks + rank) * 4;
            slot[0] = s_;
            slot[1] = c_;
            __threadfence_system();
            slot[2] = a.epoch;
        }
        if (!a.step_fused) return;
        // stop test (k_sharded_step / k_sharded_step_fused as an epilogue: no second launch per level)
        if (tid < 32) {
            bool ok = true;
            unsigned long long waited = 0ULL;
            const double* mail = a.peer_mail[a.rank];
            if (exchange) ok = peer_wait_flags(mail, slot_row, nranks, a.epoch, a.timeout_ns, &waited);
            if (tid == 0) {
                if (!ok) a.state_out[5] = num<T>::from_double(3.0);
                else if (exchange)
                    level_step_apply<T, D>(a.low, a.up, a.tm, a.rtol, a.atol, level, nranks,
                                           [&](int r) {
                                               const volatile double* slot = mail + ((size_t)slot_row * nranks + r) * 4;
                                               Comp<T> o;
                                               o.s = num<T>::from_double(slot[0]);
                                               o.c = num<T>::from_double(slot[1]);
                                               return o;
                                           },
                                           a.state_out, a.options, a.max_levels);
                else
                    level_step_apply<T, D>(a.low, a.up, a.tm, a.rtol, a.atol, level, 1, [&](int) { return tot; }, a.state_out, a.options, a.max_levels);
                if (a.wait_ns) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    a.wait_ns[level] = waited;
                    if (level < 32) a.wait_ns[32 + level] = now;
                }
                if (a.gen) {
                    s_gen = ((a.gen_base + (unsigned int)(level + 1)) << 1) | (a.state_out[5] != num<T>::zero() ? 1u : 0u);
                    st_release_gpu(a.gen, s_gen);
                } else {
                    __threadfence();
                }
            }
        }
        __syncthreads();  // the next level re-stages shared memory and reads the state
    }
}

// Stop test of the two-launch fused path (FTS_B200_STEP_FUSED=0; the default does this in the epilogue of
// k_level_grid): lane r of one warp waits for rank r's flag in OUR mailbox (peers write it through NVLink),
// then thread 0 does the arithmetic of k_sharded_step.  A rank that never arrives trips the timeout
// (FTS_B200_PEER_TIMEOUT_MS, default 10 s: host-side skew between ranks — first-call module loads, a garbage
// collection — must not look like a dead peer) -> state[5] = 3 and every later launch of the call exits.
template <class T, int D> __global__ void k_sharded_step_fused(const T* low, const T* up, T tm, T rtol, T atol, int level, int nranks,
                                                               const double* mail, int mail_slot, double epoch, T* state, int options,
                                                               unsigned long long timeout_ns, unsigned long long* wait_ns, int max_levels)
{
    if (level > 0 && state[5] != num<T>::zero()) {
        if (wait_ns && threadIdx.x == 0) wait_ns[level] = 0ULL;
        return;
    }
    // time this rank spent waiting for the slowest peer's flag (0 when every flag was already there):
    // the per-level exchange wait bench.py reports for the sharded 3D integral
    unsigned long long waited = 0ULL;
    const bool ok = peer_wait_flags(mail, mail_slot, nranks, epoch, timeout_ns, &waited);
    if (threadIdx.x != 0) return;
    if (wait_ns) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        wait_ns[level] = waited;
        if (level < 32) wait_ns[32 + level] = now;
    }
    if (!ok) {
        state[5] = num<T>::from_double(3.0);
        return;
    }
    level_step_apply<T, D>(low, up, tm, rtol, atol, level, nranks,
                           [&](int r) {
                               const volatile double* slot = mail + ((size_t)mail_slot * nranks + r) * 4;
                               Comp<T> o;
                               o.s = num<T>::from_double(slot[0]);
                               o.c = num<T>::from_double(slot[1]);
                               return o;
                           },
                           state, options, max_levels);
}

// Stop test with the partials of all ranks in device memory (the caller exchanged them: NCCL all-gather, or
// one rank): partials = [nranks][2] (s, c) in rank order.
template <class T, int D> __global__ void k_sharded_step(const T* low, const T* up, T tm, T rtol, T atol, int level, int nranks,
                                                         const T* partials, T* state, int options, int max_levels)
{
    if (level > 0 && state[5] != num<T>::zero()) return;
    level_step_apply<T, D>(low, up, tm, rtol, atol, level, nranks,
                           [&](int r) {
                               Comp<T> o;
                               o.s = partials[2 * r];
                               o.c = partials[2 * r + 1];
                               return o;
                           },
                           state, options, max_levels);
}

}  // namespace fts
