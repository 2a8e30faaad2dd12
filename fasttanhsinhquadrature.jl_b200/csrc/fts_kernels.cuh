// fts_kernels.cuh — sm_100a kernels of the tanh-sinh hot path (fixed grids + adaptive levels).
//
// Two kernel styles per problem class:
//   *_tpi  thread-per-integral, FTS_ORDER_REFERENCE: the literal sequential summation order and
//          un-contracted arithmetic of the reference's scalar kernels (what quad/quad_cmpl/
//          integrate*D compute).  Node tables are read with warp-uniform 128-bit loads
//          (all lanes of a warp walk the same level), per-integral data is coalesced.
//   *_cta  CTA-per-integral, FTS_ORDER_FAST: nodes of a level are dealt to the threads of the
//          CTA, FMA accumulation, compensated (TwoSum) warp-shuffle + shared-memory merge,
//          on-device stop test (the analogue of the reference's @turbo `_avx` twins).
//
// Every kernel takes ONE POD argument block (fts::Args<T>) so that statically compiled builtin
// families and NVRTC-compiled user integrands are launched through the same code path.
// This header is fed verbatim to NVRTC: no host includes.
#pragma once
#include "fts_math.cuh"

namespace fts {

enum : int { OPT_QUAD_EDGE = 1, OPT_FORCE_DEPTH = 2 };
// FTS_OPT_FORCE_DEPTH: the stop rule only counts at the last level
__device__ __forceinline__ bool may_stop(int options, int level, int max_levels) { return !(options & OPT_FORCE_DEPTH) || level == max_levels; }
enum : int { FLAG_CONVERGED = 1, FLAG_MAX_LEVELS = 2, FLAG_ZERO_WIDTH = 4, FLAG_FLIPPED = 8 };

#ifndef FTS_MAX_RANKS
#define FTS_MAX_RANKS 16
#endif
#ifndef FTS_TPI_PAIRS
#define FTS_TPI_PAIRS 4
#endif
// trips of that loop per branch in the kernels whose evaluation is short (table exp without range test: EXP_SAFE):
// loop control and the constant reloads are 11 of 167 instructions per trip there (config B 0.2682 -> 0.2659 ms)
#ifndef FTS_TPI_UNROLL
#define FTS_TPI_UNROLL 4
#endif
constexpr int TPI_PAIRS = FTS_TPI_PAIRS;  // node pairs per loop trip of the thread-per-integral 1D kernels (levels >= 2)

template <class T> struct Node { T x, w; };             // 1D / tensor tables, interleaved
template <class T> struct NodeC { T x, w, c, pad; };    // complement tables

struct PeerPtrs { double* p[FTS_MAX_RANKS]; };  // the ranks' mailboxes as a kernel argument

template <class T> struct Args {
    // fixed grid (x,w,h) = tanhsinh(T,N)
    const Node<T>* grid;
    int n;
    T h;
    // adaptive level tables (flat, level offsets implied: 1D 2^l-2, tensor 2^(l+1)-4)
    const Node<T>* nodes;
    const NodeC<T>* nodes_c;
    T tm;
    T ix[2], iw[2], ic[2];
    // per-integral inputs
    const T* low;
    const T* up;
    int bstride;
    const T* params;
    int pstride;
    T rtol, atol;
    int max_levels;
    int options;
    long long batch;
    // outputs
    T* value;
    T* err;
    int* levels;
    int* flags;
    // grid-per-level kernels (one large integral, optionally sharded over ranks)
    int level;            // level computed by this launch (0 = base grid)
    int rank, nranks;     // shard of the flat new-cell index handled by this device
    T* block_partials;    // [gridDim.x][2] compensated per-CTA partials (s, c)
    unsigned int* ticket; // last-CTA election
    T* partial_out;       // [2] this rank's (s, c) of the level
    const T* state;       // [8] see fts_adaptive_sharded_step_dev; state[5] != 0 -> already converged
    // deep-level hand-off of the thread-per-integral 1D kernels: an integral that has not met its
    // tolerance after level `tail_level` is queued (its running sum parked in value[b]) and finished
    // by k_adaptive1d_tail, one CTA per queued integral, in the same summation order
    int tail_level;       // < 0: disabled
    int* tail_count;
    int* tail_idx;
    // bucket passes of the 1D thread-per-integral call (k_bucket_sort, k_adaptive1d_bucket): a warp whose lanes
    // would look up unrelated exp-table rows and stop at different levels (parameters in arbitrary order) does
    // not compute; it leaves the float bits of its lanes' exp-argument bounds in sort_idx[b] and raises its bit
    // sort_mask(b >> 5).  sort_ctl = {any warp queued, next warp task of k_adaptive1d_bucket}.
    // sort_idx = nullptr: not armed.  Layout behind sort_idx (B = batch rounded up to SORT_CHUNK):
    // keys[B], indices in bucket order per chunk (-1 behind the queued ones)[B], queued-warp flags[B / 32]
    int* sort_ctl;
    int* sort_idx;
    int* sort_fb;         // pinned host word: the tail kernel stores sort_ctl[0] there (feedback for the next call); may be null
    // launch shape of the grouped 2D adaptive kernel (k_adaptive2d_cta): 0 = CTA per integral,
    // 32 = warp per integral with `level` as the deepest level a warp handles
    int group;
    // fused exchange over NVLink peer memory (k_level_grid epilogue): the last CTA stores this
    // rank's (s, c) and then the epoch flag straight into every rank's mailbox
    double* peer_mail[FTS_MAX_RANKS];  // mailbox of rank r mapped into this process (own one included); null: no exchange
    int mail_slot;                     // index of this launch's slot row: (epoch&1)*levels + level (of a.level; +1 per level of a range)
    double epoch;                      // flag value of this call
    // stop test in the epilogue of k_level_grid (step_fused != 0): the last CTA of the launch waits for the peers'
    // partials (if any), updates state_out like k_sharded_step and leaves; level_hi > level: one CTA walks the
    // levels level ... level_hi in one launch (the levels that are one CTA small)
    int step_fused;
    int level_hi;
    int shard_from;                    // levels of the range below this one: whole on every rank, no exchange
    unsigned int* gen;                 // generation barrier of a multi-CTA range (cooperative launch); null: one CTA or one level
    unsigned short range_blocks[32];   // CTAs that work on level l of a multi-CTA range (<= gridDim.x)
    unsigned int gen_base;             // the barrier counts gen_base + level + 1 (one zeroed word serves the integrals of a call one after the other)
    T* state_out;
    unsigned long long timeout_ns;     // peer wait limit
    unsigned long long* wait_ns;       // [64] or null: [level] exchange wait, [32 + level] globaltimer when the level's stop test was done
};

template <int NP, class T> __device__ __forceinline__ void load_params(T (&p)[NP > 0 ? NP : 1], const Args<T>& a, long long b)
{
#pragma unroll
    for (int k = 0; k < NP; ++k) p[k] = a.params[b * a.pstride + k];
    if (NP == 0) p[0] = num<T>::zero();
}

// Integrand evaluation at N independent points.  Families whose cost is one long dependent FP64
// chain (exp) provide `evaln`, which interleaves the N chains (fts_math.cuh exp_n); every other
// functor — NVRTC user bodies included — falls back to N scalar evaluations.
template <class F, class = void> struct HasEvalN { static constexpr bool value = false; };
template <class F> struct HasEvalN<F, decltype((void)F::HAS_EVALN)> { static constexpr bool value = F::HAS_EVALN; };
// Families whose Float64 code reads a lookup table from SHARED memory declare it (USES_EXP_TAB:
// fts_math.cuh exp_n with EXP_SMEM; USES_LOG_TAB: log_n); every kernel copies the tables it needs at
// entry — family_tables_init contains a barrier, so it runs before any thread can leave.
template <class T> struct IsF64 { static constexpr bool value = false; };
template <> struct IsF64<double> { static constexpr bool value = true; };
template <class F, class = void> struct DeclaresExpTab { static constexpr bool value = false; };
template <class F> struct DeclaresExpTab<F, decltype((void)F::USES_EXP_TAB)> { static constexpr bool value = F::USES_EXP_TAB; };
template <class F, class = void> struct DeclaresLogTab { static constexpr bool value = false; };
template <class F> struct DeclaresLogTab<F, decltype((void)F::USES_LOG_TAB)> { static constexpr bool value = F::USES_LOG_TAB; };
template <class T, class F> struct UsesExpTab { static constexpr bool value = DeclaresExpTab<F>::value && IsF64<T>::value; };
template <class T, class F> struct UsesLogTab { static constexpr bool value = DeclaresLogTab<F>::value && IsF64<T>::value; };
// programmatic dependent launch (see launch_dependent in fts_host.cu): the first kernel of a two-kernel call
// lets its dependent be scheduled early, the dependent waits for the first kernel's completion and memory
// flush before it reads the queue.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifndef FTS_SORT_CHUNK
#define FTS_SORT_CHUNK 1024
#endif
constexpr int SORT_CHUNK = FTS_SORT_CHUNK;  // integrals one CTA of k_bucket_sort orders (a multiple of 256)
template <class T> __device__ __forceinline__ long long sort_padded(const Args<T>& a) { return (a.batch + SORT_CHUNK - 1) / SORT_CHUNK * SORT_CHUNK; }
template <class T> __device__ __forceinline__ int* sort_sorted(const Args<T>& a) { return a.sort_idx + sort_padded(a); }
template <class T> __device__ __forceinline__ int* sort_mask(const Args<T>& a) { return a.sort_idx + 2 * sort_padded(a); }

template <class T, class F> __device__ __forceinline__ void family_tables_init()
{
    if constexpr (UsesExpTab<T, F>::value) exp_tab_init();
    if constexpr (UsesLogTab<T, F>::value) log_tab_init();
}

template <class T, class F, bool FAST, int N, int MODE = 0> __device__ __forceinline__ void eval1d_n(const T (&x)[N], const T* p, T (&f)[N])
{
    if constexpr (HasEvalN<F>::value) {
        F::template evaln<T, FAST, N, MODE>(x, p, f);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = F::template eval<T, FAST>(x[i], p);
    }
}

template <class T, class F, bool FAST, int N, int MODE = 0>
__device__ __forceinline__ void eval2d_n(const T (&x)[N], const T (&y)[N], const T* p, T (&f)[N])
{
    if constexpr (HasEvalN<F>::value) {
        F::template evaln<T, FAST, N, MODE>(x, y, p, f);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = F::template eval<T, FAST>(x[i], y[i], p);
    }
}
template <class T, class F, bool FAST, int N, int MODE = 0>
__device__ __forceinline__ void eval3d_n(const T (&x)[N], const T (&y)[N], const T (&z)[N], const T* p, T (&f)[N])
{
    if constexpr (HasEvalN<F>::value) {
        F::template evaln<T, FAST, N, MODE>(x, y, z, p, f);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = F::template eval<T, FAST>(x[i], y[i], z[i], p);
    }
}

// 2D complement protocol: f(x, y, bmx, xma, dmy, ymc)
template <class T, class F, bool FAST, int N>
__device__ __forceinline__ void eval2dc_n(const T (&x)[N], const T (&y)[N], const T (&bmx)[N], const T (&xma)[N], const T (&dmy)[N], const T (&ymc)[N],
                                          const T* p, T (&f)[N])
{
    if constexpr (HasEvalN<F>::value) {
        F::template evaln<T, FAST, N, 0>(x, y, bmx, xma, dmy, ymc, p, f);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = F::template eval<T, FAST>(x[i], y[i], bmx[i], xma[i], dmy[i], ymc[i], p);
    }
}
// families that evaluate a run of cells along k with the work shared between them (see F3Log::cells_run)
template <class F, class = void> struct HasCellsRun { static constexpr bool value = false; };
template <class F> struct HasCellsRun<F, decltype((void)F::HAS_CELLS_RUN)> { static constexpr bool value = F::HAS_CELLS_RUN; };

template <class T> __device__ __forceinline__ Node<T> ldg_node(const Node<T>* p) { return *p; }
template <> __device__ __forceinline__ Node<double> ldg_node<double>(const Node<double>* p)
{
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    Node<double> r;
    r.x = v.x;
    r.w = v.y;
    return r;
}
template <> __device__ __forceinline__ Node<float> ldg_node<float>(const Node<float>* p)
{
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    Node<float> r;
    r.x = v.x;
    r.w = v.y;
    return r;
}

template <class T> __device__ __forceinline__ void store_result(const Args<T>& a, long long b, T value, T err, int level, int flags)
{
    a.value[b] = value;
    if (a.err) a.err[b] = err;
    if (a.levels) a.levels[b] = level;
    if (a.flags) a.flags[b] = flags;
}

// quad()/quad_cmpl() edge rules for one axis (quad.jl:41-46,303-308): returns false when the
// integral is identically zero; flips bounds and toggles `neg` when low > up.
template <class T> __device__ __forceinline__ bool quad_edge_axis(T& lo, T& hi, bool& neg)
{
    if (lo == hi) return false;
    if (lo > hi) {
        T t = lo;
        lo = hi;
        hi = t;
        neg = !neg;
    }
    return true;
}

// Pointwise values of a 1D integrand, f(x_b) with FMA contraction (FAST arithmetic): lets a caller
// check a builtin family or an NVRTC body on the device (tests hold the table-driven exp/log to their
// ulp bounds with it).  x is passed in `low`.
template <class T, class F> __global__ void __launch_bounds__(256) k_eval1d(const Args<T> a)
{
    family_tables_init<T, F>();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    a.value[b] = F::template eval<T, true>(a.low[b], p);
}

// =================================================================== fixed grids, reference order
// integrate1D(f, low, up, x, w, h)   /root/reference/src/integrate1D.jl:90-102
template <class T, class F, bool FAST> __global__ void __launch_bounds__(256) k_fixed1d_tpi(const Args<T> a)
{
    constexpr int MODE = UsesExpTab<T, F>::value ? EXP_SMEM : 0;
    family_tables_init<T, F>();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T lo = a.low[b * a.bstride], hi = a.up[b * a.bstride];
    const T dx = half * (hi - lo), x0 = half * (hi + lo);  // core.jl:27-30
    T s = num<T>::half_pi() * F::template eval<T, FAST>(x0, p);
    const int n = a.n;
    constexpr int PAIRS = TPI_PAIRS;
    int i = 0;
#pragma unroll 1
    for (; i + PAIRS <= n; i += PAIRS) {  // PAIRS node pairs = 2*PAIRS independent evaluations per trip, summed in order
        T xs[2 * PAIRS], ws[PAIRS], fs[2 * PAIRS];
#pragma unroll
        for (int k = 0; k < PAIRS; ++k) {
            const Node<T> nd = ldg_node(a.grid + i + k);
            const T d = dx * nd.x;
            xs[2 * k] = x0 - d;      // f(xm) + f(xp) in this order (integrate1D.jl:97-99)
            xs[2 * k + 1] = x0 + d;
            ws[k] = nd.w;
        }
        eval1d_n<T, F, FAST, 2 * PAIRS, MODE>(xs, p, fs);
#pragma unroll
        for (int k = 0; k < PAIRS; ++k) s = muladd<T, FAST>(ws[k], fs[2 * k] + fs[2 * k + 1], s);
    }
    for (; i < n; ++i) {
        const Node<T> nd = ldg_node(a.grid + i);
        const T d = dx * nd.x;
        const T xp = x0 + d, xm = x0 - d;
        s = muladd<T, FAST>(nd.w, F::template eval<T, FAST>(xm, p) + F::template eval<T, FAST>(xp, p), s);
    }
    a.value[b] = dx * a.h * s;
}

// integrate1D_cmpl(T, f, N)   /root/reference/src/integrate1D.jl:49-65: f(x, 1-x, 1+x) over [-1, 1] on the
// fixed grid, the complements formed by subtraction (not complement-accurate, App. B.9); one thread
// per parameter row
template <class T, class F, bool FAST> __global__ void __launch_bounds__(256) k_fixed1d_cmpl_tpi(const Args<T> a)
{
    family_tables_init<T, F>();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T one = num<T>::one();
    T s = num<T>::half_pi() * F::template eval<T, FAST>(num<T>::zero(), one, one, p);
    for (int i = 0; i < a.n; ++i) {
        const Node<T> nd = ldg_node(a.grid + i);
        const T omx = one - nd.x, opx = one + nd.x;
        s = muladd<T, FAST>(nd.w, F::template eval<T, FAST>(nd.x, omx, opx, p) + F::template eval<T, FAST>(-nd.x, opx, omx, p), s);
    }
    a.value[b] = a.h * s;
}

// integrate2D(f, low, up, x, w, h)   /root/reference/src/integrate2D.jl:91-124
template <class T, class F, bool FAST> __global__ void __launch_bounds__(128) k_fixed2d_tpi(const Args<T> a)
{
    family_tables_init<T, F>();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T* lo = a.low + b * a.bstride;
    const T* hi = a.up + b * a.bstride;
    const T dx = half * (hi[0] - lo[0]), x0 = half * (hi[0] + lo[0]);
    const T dy = half * (hi[1] - lo[1]), y0 = half * (hi[1] + lo[1]);
    const T w0 = num<T>::half_pi();
    const T w02 = w0 * w0;
    T s = w02 * F::template eval<T, FAST>(x0, y0, p);
    const int n = a.n;
    for (int k = 0; k < n; ++k) {
        const Node<T> nd = ldg_node(a.grid + k);
        const T xkp = dx * nd.x + x0, xkm = x0 - dx * nd.x;
        const T ykp = dy * nd.x + y0, ykm = y0 - dy * nd.x;
        s = s + nd.w * w0 * (F::template eval<T, FAST>(xkp, y0, p) + F::template eval<T, FAST>(xkm, y0, p) +
                             F::template eval<T, FAST>(x0, ykp, p) + F::template eval<T, FAST>(x0, ykm, p));
    }
    for (int i = 0; i < n; ++i) {
        const Node<T> ni = ldg_node(a.grid + i);
        const T xip = dx * ni.x + x0, xim = x0 - dx * ni.x;
        T inner = num<T>::zero();
        for (int j = 0; j < n; ++j) {
            const Node<T> nj = ldg_node(a.grid + j);
            const T yjp = dy * nj.x + y0, yjm = y0 - dy * nj.x;
            inner = inner + nj.w * (F::template eval<T, FAST>(xim, yjm, p) + F::template eval<T, FAST>(xip, yjm, p) +
                                    F::template eval<T, FAST>(xim, yjp, p) + F::template eval<T, FAST>(xip, yjp, p));
        }
        s = s + ni.w * inner;
    }
    a.value[b] = dx * dy * (a.h * a.h) * s;
}

// integrate3D(f, low, up, x, w, h)   /root/reference/src/integrate3D.jl:150-207
template <class T, class F, bool FAST> __global__ void __launch_bounds__(128) k_fixed3d_tpi(const Args<T> a)
{
    family_tables_init<T, F>();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T* lo = a.low + b * a.bstride;
    const T* hi = a.up + b * a.bstride;
    const T dx = half * (hi[0] - lo[0]), x0 = half * (hi[0] + lo[0]);
    const T dy = half * (hi[1] - lo[1]), y0 = half * (hi[1] + lo[1]);
    const T dz = half * (hi[2] - lo[2]), z0 = half * (hi[2] + lo[2]);
    const T w0 = num<T>::half_pi();
    const T w02 = w0 * w0;
    const T w03 = w02 * w0;
#define FE(X, Y, Z) F::template eval<T, FAST>(X, Y, Z, p)
    T total = w03 * FE(x0, y0, z0);
    const int n = a.n;
    for (int i = 0; i < n; ++i) {
        const Node<T> nd = ldg_node(a.grid + i);
        const T axis = (FE(dx * nd.x + x0, y0, z0) + FE(x0 - dx * nd.x, y0, z0)) + (FE(x0, dy * nd.x + y0, z0) + FE(x0, y0 - dy * nd.x, z0)) +
                       (FE(x0, y0, dz * nd.x + z0) + FE(x0, y0, z0 - dz * nd.x));
        total = total + nd.w * w02 * axis;
    }
    for (int i = 0; i < n; ++i) {
        const Node<T> ni = ldg_node(a.grid + i);
        const T xp = dx * ni.x + x0, xm = x0 - dx * ni.x;
        const T ypi = dy * ni.x + y0, ymi = y0 - dy * ni.x;
        for (int j = 0; j < n; ++j) {
            const Node<T> nj = ldg_node(a.grid + j);
            const T wiwj = ni.w * nj.w;
            const T yp = dy * nj.x + y0, ym = y0 - dy * nj.x;
            const T zpj = dz * nj.x + z0, zmj = z0 - dz * nj.x;
            const T plane = (FE(xp, yp, z0) + FE(xm, yp, z0) + FE(xp, ym, z0) + FE(xm, ym, z0)) +
                            (FE(xp, y0, zpj) + FE(xm, y0, zpj) + FE(xp, y0, zmj) + FE(xm, y0, zmj)) +
                            (FE(x0, ypi, zpj) + FE(x0, ymi, zpj) + FE(x0, ypi, zmj) + FE(x0, ymi, zmj));
            total = total + wiwj * w0 * plane;
            T oct = num<T>::zero();
            for (int k = 0; k < n; ++k) {
                const Node<T> nk = ldg_node(a.grid + k);
                const T zp = dz * nk.x + z0, zm = z0 - dz * nk.x;
                oct = oct + nk.w * ((FE(xp, yp, zp) + FE(xm, yp, zp) + FE(xp, ym, zp) + FE(xm, ym, zp)) +
                                    (FE(xp, yp, zm) + FE(xm, yp, zm) + FE(xp, ym, zm) + FE(xm, ym, zm)));
            }
            total = total + wiwj * oct;
        }
    }
#undef FE
    a.value[b] = (dx * dy * dz * (a.h * a.h * a.h)) * total;
}

// =================================================================== adaptive, reference order
// adaptive_integrate_1D typed core   /root/reference/src/adaptive.jl:28-75

// levels of one integral; MODE selects the exp variant (shared-memory table, no range test)
template <class T, class F, bool FAST, int MODE>
__device__ __forceinline__ void adaptive1d_tpi_levels(const Args<T>& a, long long b, const T* p, T dx, T x0, bool neg, int flags, int* q_count, int* q_idx)
{
    const T half = num<T>::half();
    T h = a.tm * half;
    const T w0 = num<T>::pi() * half;
    T s_total;
    {   // level 0 (adaptive.jl:39-45): centre + the two initial node pairs, 5 evaluations in flight
        const T d0 = dx * a.ix[0], d1 = dx * a.ix[1];
        const T xs[5] = {x0, d0 + x0, x0 - d0, d1 + x0, x0 - d1};
        T fs[5];
        eval1d_n<T, F, FAST, 5, MODE>(xs, p, fs);
        s_total = w0 * fs[0];
        s_total = s_total + a.iw[0] * (fs[1] + fs[2]);
        s_total = s_total + a.iw[1] * (fs[3] + fs[4]);
    }
    T old_res = dx * h * s_total;
    T err = num<T>::zero();
    int level_used = 0;
    const Node<T>* __restrict__ nodes = a.nodes;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        T s_new = num<T>::zero();
        const int n = 1 << level;
        const Node<T>* __restrict__ lv = nodes + (n - 2);
        // PAIRS node pairs per trip (n = 2^level): 2*PAIRS independent evaluations in flight,
        // accumulated in the reference's order.  Level 1 has only two pairs.
        if (level == 1) {
            const Node<T> n0 = ldg_node(lv), n1 = ldg_node(lv + 1);
            const T d0 = dx * n0.x, d1 = dx * n1.x;
            const T xs[4] = {d0 + x0, x0 - d0, d1 + x0, x0 - d1};
            T fs[4];
            eval1d_n<T, F, FAST, 4, MODE>(xs, p, fs);
            s_new = muladd<T, FAST>(n0.w, fs[0] + fs[1], s_new);
            s_new = muladd<T, FAST>(n1.w, fs[2] + fs[3], s_new);
        } else {
            constexpr int PAIRS = TPI_PAIRS;
            auto trip = [&](int i) {
                T xs[2 * PAIRS], ws[PAIRS], fs[2 * PAIRS];
#pragma unroll
                for (int k = 0; k < PAIRS; ++k) {
                    const Node<T> nd = ldg_node(lv + i + k);
                    const T d = dx * nd.x;
                    xs[2 * k] = d + x0;
                    xs[2 * k + 1] = x0 - d;
                    ws[k] = nd.w;
                }
                eval1d_n<T, F, FAST, 2 * PAIRS, MODE>(xs, p, fs);
#pragma unroll
                for (int k = 0; k < PAIRS; ++k) s_new = muladd<T, FAST>(ws[k], fs[2 * k] + fs[2 * k + 1], s_new);
            };
            constexpr int UNR = (MODE & EXP_SAFE) ? FTS_TPI_UNROLL : 1;
            if (UNR > 1 && n >= UNR * PAIRS) {
#pragma unroll 1
                for (int i = 0; i < n; i += UNR * PAIRS) {
#pragma unroll
                    for (int u = 0; u < UNR; ++u) trip(i + u * PAIRS);
                }
            } else {
#pragma unroll 1
                for (int i = 0; i < n; i += PAIRS) trip(i);
            }
        }
        s_total = s_total + s_new;
        const T new_res = dx * h * s_total;
        err = fm::abs(new_res - old_res);
        level_used = level;
        old_res = new_res;
        if (err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels)) {
            flags |= FLAG_CONVERGED;
            break;
        }
        if (level == a.tail_level && level < a.max_levels) {  // park and let a whole CTA finish it
            q_idx[atomicAdd(q_count, 1)] = (int)b;
            a.value[b] = s_total;
            return;
        }
    }
    if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
    store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
}

// quad() edge rules and the affine map of integral b with bounds (lo, hi); false: the result (zero width) has been stored
template <class T, class F> __device__ __forceinline__ bool tpi_finish_prologue(const Args<T>& a, long long b, T lo, T hi, T& dx, T& x0, bool& neg, int& flags)
{
    const T half = num<T>::half();
    neg = false;
    flags = 0;
    if (a.options & OPT_QUAD_EDGE) {
        if (!quad_edge_axis(lo, hi, neg)) {
            store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
            return false;
        }
        if (neg) flags |= FLAG_FLIPPED;
    }
    dx = half * (hi - lo);
    x0 = half * (hi + lo);
    return true;
}

#ifndef FTS_TPI_LB_THREADS
#define FTS_TPI_LB_THREADS 256
#define FTS_TPI_LB_BLOCKS 2
#endif
template <class T, class F, bool FAST> __global__ void __launch_bounds__(FTS_TPI_LB_THREADS, FTS_TPI_LB_BLOCKS) k_adaptive1d_tpi(const Args<T> a)
{
    pdl_launch_dependents();     // the kernels behind this one may be set up behind this grid's last wave
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = b < a.batch;
    T p[F::NP > 0 ? F::NP : 1];
    T lo = num<T>::zero(), hi = num<T>::zero();
    if (mine) {                  // the loads of this integral's inputs fly while the CTA fills its tables
        load_params<F::NP>(p, a, b);
        lo = a.low[b * a.bstride];
        hi = a.up[b * a.bstride];
    }
    if constexpr (UsesExpTab<T, F>::value) {
        // Bucket passes armed (uniform over the grid): a CTA whose warps ALL hand their integrals over leaves
        // before the table fill — in a batch in arbitrary order that is every CTA, and this kernel shrinks to
        // reading the parameters and writing the sort keys.  (Same classification as below; a zero-width
        // integral is in no hurry: it goes through the regular path of whichever kernel gets it.)
        if (a.sort_idx != nullptr) {
            bool queued = false;
            float key = 0.0f;
            if (mine) {
                const T half = num<T>::half();
                T l2 = lo, h2 = hi;
                bool n2 = false;
                if (a.options & OPT_QUAD_EDGE) quad_edge_axis(l2, h2, n2);
                const T bound = F::template exp_arg_bound<T>(fm::abs(half * (h2 + l2)) + fm::abs(half * (h2 - l2)), p);
                const int q = __double2int_rd(fmin((double)bound, 1.0e7) * 8.0);
                const unsigned m = __activemask();
                queued = __reduce_max_sync(m, q) - __reduce_min_sync(m, q) > 2;
                key = fminf(fmaxf((float)bound, 0.0f), 3.0e38f);
            }
            if (!__syncthreads_or(mine && !queued)) {
                if (mine) {
                    if ((threadIdx.x & 31) == 0) {
                        a.sort_ctl[0] = 1;
                        if (((b >> 5) & 7) == 0) atomicAdd(a.sort_ctl + 2, 8);  // one warp in 8 counts for 8
                        sort_mask(a)[b >> 5] = 1;
                    }
                    a.sort_idx[b] = __float_as_int(key);
                }
                return;
            }
        }
    }
    family_tables_init<T, F>();  // before any thread leaves (contains a barrier)
    if (!mine) return;
    T dx, x0;
    bool neg;
    int flags;
    if (!tpi_finish_prologue<T, F>(a, b, lo, hi, dx, x0, neg, flags)) return;
    if constexpr (UsesExpTab<T, F>::value) {
        // every exp argument of this integral is bounded by F::exp_arg_bound(|x0|+|dx|): below 700
        // the overflow/underflow test of exp is dead weight
        const T bound = F::template exp_arg_bound<T>(fm::abs(x0) + fm::abs(dx), p);
        if (a.sort_fb != nullptr) {
            // Lanes whose bounds differ by more than ~1/4 look up unrelated rows of the shared exp table (an
            // argument step of 1/4 is 46 rows: ~10 bank-conflict wavefronts per lookup instead of 1-2) and
            // stop at different levels.  With the bucket passes armed (sort_idx) such a warp does not compute: it
            // hands its integrals to them, which regroup them by bound; otherwise it only reports itself, so that
            // the next call arms the passes.  A sorted sweep never gets here.
            const int q = __double2int_rd(fmin((double)bound, 1.0e7) * 8.0);
            const unsigned m = __activemask();
            const bool queued = __reduce_max_sync(m, q) - __reduce_min_sync(m, q) > 2;
            const bool leader = (int)(threadIdx.x & 31) == __ffs(m) - 1;
            if (queued && leader) {
                a.sort_ctl[0] = 1;
                if (((b >> 5) & 7) == 0) atomicAdd(a.sort_ctl + 2, 8);  // sampled count of incoherent warps: the next call's choice of pass
            }
            if (a.sort_idx != nullptr) {
                if (leader) sort_mask(a)[b >> 5] = queued ? 1 : 0;
                if (queued) {
                    a.sort_idx[b] = __float_as_int(fminf(fmaxf((float)bound, 0.0f), 3.0e38f));
                    return;
                }
            }
        }
        if (bound < 700.0)
            adaptive1d_tpi_levels<T, F, FAST, EXP_SMEM | EXP_SAFE>(a, b, p, dx, x0, neg, flags, a.tail_count, a.tail_idx);
        else
            adaptive1d_tpi_levels<T, F, FAST, EXP_SMEM>(a, b, p, dx, x0, neg, flags, a.tail_count, a.tail_idx);
    } else {
        adaptive1d_tpi_levels<T, F, FAST, 0>(a, b, p, dx, x0, neg, flags, a.tail_count, a.tail_idx);
    }
}

// adaptive_integrate_1D_cmpl typed core   /root/reference/src/adaptive.jl:727-785
template <class T, class F, bool FAST> __global__ void __launch_bounds__(256) k_adaptive1d_cmpl_tpi(const Args<T> a)
{
    pdl_launch_dependents();
    family_tables_init<T, F>();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half(), one = num<T>::one();
    T lo = a.low[b * a.bstride], hi = a.up[b * a.bstride];
    bool neg = false;
    int flags = 0;
    if (a.options & OPT_QUAD_EDGE) {
        if (!quad_edge_axis(lo, hi, neg)) {
            store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
            return;
        }
        if (neg) flags |= FLAG_FLIPPED;
    }
    const T dx = half * (hi - lo), x0 = half * (hi + lo);
    T h = a.tm * half;
    const T w0 = num<T>::half_pi();
    T s_total = w0 * F::template eval<T, FAST>(x0, dx, dx, p);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T xk = a.ix[i], ck = a.ic[i], wk = a.iw[i];
        const T dxck = dx * ck;
        const T dx1pxk = dx * (one + xk);
        s_total = s_total + wk * (F::template eval<T, FAST>(dx * xk + x0, dxck, dx1pxk, p) + F::template eval<T, FAST>(x0 - dx * xk, dx1pxk, dxck, p));
    }
    T old_res = dx * h * s_total;
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        T s_new = num<T>::zero();
        const int n = 1 << level;
        const NodeC<T>* __restrict__ lv = a.nodes_c + (n - 2);
        for (int i = 0; i < n; ++i) {
            const NodeC<T> nd = lv[i];
            const T dxck = dx * nd.c;
            const T dx1pxk = dx * (one + nd.x);
            const T d = dx * nd.x;
            s_new = muladd<T, FAST>(nd.w, F::template eval<T, FAST>(d + x0, dxck, dx1pxk, p) + F::template eval<T, FAST>(x0 - d, dx1pxk, dxck, p), s_new);
        }
        s_total = s_total + s_new;
        const T new_res = dx * h * s_total;
        err = fm::abs(new_res - old_res);
        level_used = level;
        old_res = new_res;
        if (err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels)) {
            flags |= FLAG_CONVERGED;
            break;
        }
        if (level == a.tail_level && level < a.max_levels) {  // park and let a whole CTA finish it
            a.tail_idx[atomicAdd(a.tail_count, 1)] = (int)b;
            a.value[b] = s_total;
            return;
        }
    }
    if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
    store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
}

// ---- a group of G lanes per integral (Double64, 1D) -------------------------------------------------------
// A Double64 evaluation is 100-300 FP64 instructions of mostly dependent arithmetic, and BASELINE config E
// (65 536 integrals) is 3.5 warps per scheduler with one thread per integral: the FP64 pipe idles two thirds
// of the time.  Here G neighbouring lanes share an integral: lane g takes the nodes g, g + G, ... of a level,
// sums them in index order, and the G partial sums are added in lane order (a fixed tree of shuffles), so the
// result depends on G but not on the batch or the launch.  Nothing pins the summation order of Double64
// (the reference holds no Double64 value; tolerance 1e-28 against a 1e-32 format).  Integrals still
// unconverged at Args::tail_level join the deep queue like those of the thread-per-integral kernels.
template <class T> __device__ __forceinline__ T shfl_xor_t(unsigned mask, T v, int lane_mask)
{
    if constexpr (sizeof(T) == 16) {
        T r;
        r.hi = __shfl_xor_sync(mask, v.hi, lane_mask);
        r.lo = __shfl_xor_sync(mask, v.lo, lane_mask);
        return r;
    } else {
        return __shfl_xor_sync(mask, v, lane_mask);
    }
}
template <class T> __device__ __forceinline__ T shfl_t(unsigned mask, T v, int src)
{
    if constexpr (sizeof(T) == 16) {
        T r;
        r.hi = __shfl_sync(mask, v.hi, src);
        r.lo = __shfl_sync(mask, v.lo, src);
        return r;
    } else {
        return __shfl_sync(mask, v, src);
    }
}
// sum over the lanes of a group, the same bits on every lane (the group's first lane's value is broadcast)
template <class T, int G> __device__ __forceinline__ T group_sum(unsigned mask, T v, int first_lane)
{
#pragma unroll
    for (int m = 1; m < G; m <<= 1) v = v + shfl_xor_t<T>(mask, v, m);
    return shfl_t<T>(mask, v, first_lane);
}

template <class T, class F, bool CMPL, int G> __global__ void __launch_bounds__(256) k_adaptive1d_grp(const Args<T> a)
{
    static_assert(G == 2 || G == 4 || G == 8, "group sizes");
    pdl_launch_dependents();
    family_tables_init<T, F>();
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long b = gt / G;
    if (b >= a.batch) return;  // whole groups leave together (blockDim.x is a multiple of G)
    const int lane = threadIdx.x & 31, g = lane & (G - 1), first = lane - g;
    const unsigned mask = ((1u << G) - 1u) << first;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half(), one = num<T>::one();
    T lo = a.low[b * a.bstride], hi = a.up[b * a.bstride];
    bool neg = false;
    int flags = 0;
    if (a.options & OPT_QUAD_EDGE) {
        if (!quad_edge_axis(lo, hi, neg)) {
            if (g == 0) store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
            return;
        }
        if (neg) flags |= FLAG_FLIPPED;
    }
    const T dx = half * (hi - lo), x0 = half * (hi + lo);
    auto pair = [&](T xk, T ck, T wk) -> T {
        const T d = dx * xk;
        if constexpr (CMPL) {
            const T dxck = dx * ck, dx1p = dx * (one + xk);
            return wk * (F::template eval<T, false>(d + x0, dxck, dx1p, p) + F::template eval<T, false>(x0 - d, dx1p, dxck, p));
        } else {
            return wk * (F::template eval<T, false>(d + x0, p) + F::template eval<T, false>(x0 - d, p));
        }
    };
    T h = a.tm * half;
    // base grid: centre, pair 0, pair 1 on lanes 0, 1, 2
    T part = num<T>::zero();
    if (g == 0) {
        if constexpr (CMPL) part = num<T>::half_pi() * F::template eval<T, false>(x0, dx, dx, p);
        else part = num<T>::half_pi() * F::template eval<T, false>(x0, p);
    }
    if constexpr (G >= 4) {
        if (g == 1 || g == 2) part = pair(a.ix[g - 1], a.ic[g - 1], a.iw[g - 1]);
    } else {
        if (g == 1) part = pair(a.ix[0], a.ic[0], a.iw[0]);
        else part = part + pair(a.ix[1], a.ic[1], a.iw[1]);
    }
    T s_total = group_sum<T, G>(mask, part, first);
    T old_res = dx * h * s_total;
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        const int n = 1 << level;
        part = num<T>::zero();
        for (int i = g; i < n; i += G) {
            if constexpr (CMPL) {
                const NodeC<T> nd = a.nodes_c[(n - 2) + i];
                part = part + pair(nd.x, nd.c, nd.w);
            } else {
                const Node<T> nd = ldg_node(a.nodes + (n - 2) + i);
                part = part + pair(nd.x, nd.x, nd.w);
            }
        }
        s_total = s_total + group_sum<T, G>(mask, part, first);
        const T new_res = dx * h * s_total;
        err = fm::abs(new_res - old_res);
        level_used = level;
        old_res = new_res;
        if (err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels)) {
            flags |= FLAG_CONVERGED;
            break;
        }
        if (level == a.tail_level && level < a.max_levels) {  // park and let a whole CTA finish it
            if (g == 0) {
                a.tail_idx[atomicAdd(a.tail_count, 1)] = (int)b;
                a.value[b] = s_total;
            }
            return;
        }
    }
    if (g == 0) {
        if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
}

// s += terms[0] + terms[1] + ... in index order, one thread.  The chain of dependent adds is the
// whole cost, so the loads run eight terms ahead of it (a plain loop exposes the shared-memory
// latency once per unrolled group: 15-20 cycles per term instead of the add's 8).
template <class T> __device__ __forceinline__ void ordered_add(T& s, const T* terms, int m)
{
    T cur[8];
    int i = 0;
    if (m >= 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cur[k] = terms[k];
        for (; i + 16 <= m; i += 8) {
            T nxt[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) nxt[k] = terms[i + 8 + k];
#pragma unroll
            for (int k = 0; k < 8; ++k) s = s + cur[k];
#pragma unroll
            for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s = s + cur[k];
        i += 8;
    }
    for (; i < m; ++i) s = s + terms[i];
}

// one queued deep integral (see Args::tail_level), by the whole CTA
template <class T, class F, bool FAST, bool CMPL> __device__ __forceinline__ void tail_integral(const Args<T>& a, long long b, T* terms, int* s_done)
{
    constexpr int TILE = 1024;
    const int tid = threadIdx.x;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half(), one = num<T>::one();
    T lo = a.low[b * a.bstride], hi = a.up[b * a.bstride];
    bool neg = false;
    int flags = 0;
    if (a.options & OPT_QUAD_EDGE) {
        quad_edge_axis(lo, hi, neg);
        if (neg) flags |= FLAG_FLIPPED;
    }
    const T dx = half * (hi - lo), x0 = half * (hi + lo);
    T h = a.tm * half;
    for (int l = 0; l < a.tail_level; ++l) h = h * half;
    T s_total = a.value[b];
    T old_res = dx * h * s_total;
    T err = num<T>::zero();
    int level_used = a.tail_level;
    for (int level = a.tail_level + 1; level <= a.max_levels; ++level) {
        h = h * half;
        const int n = 1 << level;
        T s_new = num<T>::zero();
        for (int t0 = 0; t0 < n; t0 += TILE) {
            const int m = n - t0 < TILE ? n - t0 : TILE;
            for (int i = tid; i < m; i += blockDim.x) {
                if constexpr (CMPL) {
                    const NodeC<T> nd = a.nodes_c[(n - 2) + t0 + i];
                    const T dxck = dx * nd.c, dx1p = dx * (one + nd.x), d = dx * nd.x;
                    terms[i] = nd.w * (F::template eval<T, FAST>(d + x0, dxck, dx1p, p) + F::template eval<T, FAST>(x0 - d, dx1p, dxck, p));
                } else {
                    const Node<T> nd = ldg_node(a.nodes + (n - 2) + t0 + i);
                    const T d = dx * nd.x;
                    terms[i] = nd.w * (F::template eval<T, FAST>(d + x0, p) + F::template eval<T, FAST>(x0 - d, p));
                }
            }
            __syncthreads();
            if constexpr (sizeof(T) == 16) {
                // Double64: a dependent double-double add costs ~160 cycles, so the in-order sum of a
                // level-16 integral (131 072 terms) alone takes 10 ms.  Nothing pins the summation
                // order of Double64, the tolerance is 1e-28 against a 1e-32 format: the
                // tile is reduced by a fixed tree (per-thread strided sums, then halving) instead.
                T part = num<T>::zero();
                for (int i = tid; i < m; i += blockDim.x) part = part + terms[i];
                __syncthreads();
                terms[tid] = part;
                __syncthreads();
                for (int half_n = blockDim.x >> 1; half_n > 0; half_n >>= 1) {
                    if (tid < half_n) terms[tid] = terms[tid] + terms[tid + half_n];
                    __syncthreads();
                }
                if (tid == 0) s_new = s_new + terms[0];
            } else if (tid == 0) {
                ordered_add(s_new, terms, m);
            }
            __syncthreads();
        }
        if (tid == 0) {
            s_total = s_total + s_new;
            const T new_res = dx * h * s_total;
            err = fm::abs(new_res - old_res);
            level_used = level;
            old_res = new_res;
            const bool conv = err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels);
            if (conv) flags |= FLAG_CONVERGED;
            *s_done = conv ? 1 : 0;
        }
        __syncthreads();
        if (*s_done) break;
    }
    if (tid == 0) {
        if (!(flags & FLAG_CONVERGED)) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
    __syncthreads();
}

// ---- integrals in arbitrary parameter order --------------------------------------------------------
// Warps of k_adaptive1d_tpi whose lanes held unrelated integrals left them alone (Args::sort_idx).
// k_bucket_sort (one CTA per chunk of SORT_CHUNK integrals, any family / element type) orders the queued integrals
// of its chunk by the exp-argument bound their warp stored: a counting sort into SORT_CHUNK buckets between the
// chunk's smallest and largest bound, in shared memory.  k_adaptive1d_bucket then walks the ordered chunks with
// one WARP per task of 32 consecutive entries (tasks are claimed from a counter: no CTA barrier in the level loop, a
// warp that drew shallow integrals simply claims the next task): lanes of a warp now stop at the same level, and the
// eight-copy exp table (EXP_SMEM8) makes their lookups conflict-free although their rows still differ.  Every
// integral is computed by the same code as in the first kernel and writes to its own slot: results do not depend
// on the order of the batch, bit for bit.  Integrals that are still unconverged at Args::tail_level join the deep
// queue, which k_adaptive1d_tail (launched last) drains.
// counting sort of one chunk: thread `tid` holds the keys v[k] of the chunk's integrals tid + 256 k (on[k]: takes part)
__device__ __forceinline__ void bucket_sort_chunk(const float (&v)[SORT_CHUNK / 256], const bool (&on)[SORT_CHUNK / 256], int* sorted, long long c0, int* hist, int* sred)
{
    constexpr int PER = SORT_CHUNK / 256;
    const int tid = threadIdx.x;
    if (tid == 0) {
        sred[8] = 0x7f7fffff;           // bits of the largest float
        sred[9] = 0;
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) hist[tid + 256 * k] = 0;
    __syncthreads();
    int lmin = 0x7f7fffff, lmax = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k)
        if (on[k]) {
            lmin = min(lmin, __float_as_int(v[k]));  // the bits of a non-negative float are monotone in its value
            lmax = max(lmax, __float_as_int(v[k]));
        }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax = __reduce_max_sync(0xffffffffu, lmax);
    if ((tid & 31) == 0) {
        atomicMin(&sred[8], lmin);
        atomicMax(&sred[9], lmax);
    }
    __syncthreads();
    const float vmin = __int_as_float(sred[8]), vmax = __int_as_float(sred[9]);
    const float scale = vmax > vmin ? (float)(SORT_CHUNK - 1) / (vmax - vmin) : 0.0f;
    int bucket[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k)
        if (on[k]) {
            const int q = (int)((v[k] - vmin) * scale);
            bucket[k] = q < 0 ? 0 : (q > SORT_CHUNK - 1 ? SORT_CHUNK - 1 : q);
            atomicAdd(&hist[bucket[k]], 1);
        }
    __syncthreads();
    {   // exclusive scan of the bucket counts: PER consecutive buckets per thread
        int cnt[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            cnt[k] = hist[tid * PER + k];
            sum += cnt[k];
        }
        int inc = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, off);
            if ((tid & 31) >= off) inc += o;
        }
        if ((tid & 31) == 31) sred[tid >> 5] = inc;
        __syncthreads();
        int before = inc - sum;
        for (int w = 0; w < (tid >> 5); ++w) before += sred[w];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            hist[tid * PER + k] = before;
            before += cnt[k];
        }
        if (tid == 255) sred[10] = before;  // queued integrals of the chunk
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        if (on[k]) sorted[c0 + atomicAdd(&hist[bucket[k]], 1)] = (int)(c0 + tid + 256 * k);
        if (tid + 256 * k >= sred[10]) sorted[c0 + tid + 256 * k] = -1;  // the entries behind the queued ones: nothing to do
    }
}

template <class Tag> __global__ void __launch_bounds__(256) k_bucket_sort(const int* ctl, int* base, long long batch)
{
    __shared__ int hist[SORT_CHUNK];
    __shared__ int sred[12];
    pdl_launch_dependents();
    pdl_wait_primary();
    if (ctl[0] == 0) return;            // no warp queued anything: uniform over the grid
    const long long padded = (batch + SORT_CHUNK - 1) / SORT_CHUNK * SORT_CHUNK;
    const int* keys = base;
    int* sorted = base + padded;
    const int* mask = base + 2 * padded;
    const int tid = threadIdx.x;
    const long long c0 = (long long)blockIdx.x * SORT_CHUNK;
    constexpr int PER = SORT_CHUNK / 256;
    float v[PER];
    bool on[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const long long i = c0 + tid + 256 * k;
        on[k] = i < batch && mask[i >> 5] != 0;
        v[k] = on[k] ? fminf(fmaxf(__int_as_float(keys[i]), 0.0f), 3.0e38f) : 0.0f;  // a lane that left before storing its key: any bucket will do
    }
    bucket_sort_chunk(v, on, sorted, c0, hist, sred);
}

// The whole batch through the bucket passes (a call whose predecessor on the stream found MOST warps incoherent): this
// kernel is the call's first one — it computes the keys itself (same bound and same warp classification as
// k_adaptive1d_tpi) and orders every integral of its chunk; k_adaptive1d_tpi, which would only have read the
// parameters and written the keys, one 128-thread CTA per 128 integrals (8192 CTA launches for 2^20 integrals:
// 26 us), is not launched at all.  It leaves the count of incoherent warps for the next call's choice.
template <class T, class F> __global__ void __launch_bounds__(256) k_bucket_sort_all(const Args<T> a)
{
    __shared__ int hist[SORT_CHUNK];
    __shared__ int sred[12];
    pdl_launch_dependents();
    pdl_wait_primary();
    const int tid = threadIdx.x;
    const long long c0 = (long long)blockIdx.x * SORT_CHUNK;
    constexpr int PER = SORT_CHUNK / 256;
    float v[PER];
    bool on[PER];
    int incoherent = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const long long i = c0 + tid + 256 * k;
        on[k] = i < a.batch;
        v[k] = 0.0f;
        int q = 0;
        if (on[k]) {
            T p[F::NP > 0 ? F::NP : 1];
            load_params<F::NP>(p, a, i);
            const T half = num<T>::half();
            T l2 = a.low[i * a.bstride], h2 = a.up[i * a.bstride];
            bool n2 = false;
            if (a.options & OPT_QUAD_EDGE) quad_edge_axis(l2, h2, n2);
            const T bound = F::template exp_arg_bound<T>(fm::abs(half * (h2 + l2)) + fm::abs(half * (h2 - l2)), p);
            q = __double2int_rd(fmin((double)bound, 1.0e7) * 8.0);
            v[k] = fminf(fmaxf((float)bound, 0.0f), 3.0e38f);
        }
        const unsigned m = __ballot_sync(0xffffffffu, on[k]);
        if (m) {
            const int qmax = __reduce_max_sync(0xffffffffu, on[k] ? q : (int)0x80000000), qmin = __reduce_min_sync(0xffffffffu, on[k] ? q : 0x7fffffff);
            if ((tid & 31) == 0 && qmax - qmin > 2) ++incoherent;
        }
    }
    if ((tid & 31) == 0) {
        if (incoherent) atomicAdd(a.sort_ctl + 2, incoherent);
        if (tid == 0) a.sort_ctl[0] = 2;  // the bucket pass has work; 2: the count in sort_ctl[2] is exact (k_adaptive1d_tail)
    }
    bucket_sort_chunk(v, on, sort_sorted(a), c0, hist, sred);
}

template <class T, class F, bool FAST> __global__ void __launch_bounds__(256, 2) k_adaptive1d_bucket(const Args<T> a)
{
    pdl_launch_dependents();
    pdl_wait_primary();
    if (a.sort_ctl[0] == 0) return;     // uniform over the grid
    if constexpr (UsesExpTab<T, F>::value) {
        exp_tab8_init();
        const int ntasks = (int)(sort_padded(a) / 32), nchunks = (int)(sort_padded(a) / SORT_CHUNK);
        const int* sorted = sort_sorted(a);
        const int lane = threadIdx.x & 31;
        // Task t = entries 32 t ... 32 t + 31 of the ordered chunks, claimed from a counter one at a time.  (Tried and
        // dropped, profiles/r2_bucket_pass_experiments.md: claiming two tasks ahead so that the claim, the entry and the
        // inputs of a task are in flight behind the arithmetic of the previous one — every warp then sits on two tasks
        // when the counter runs out and the kernel ends 10 us later; dealing most tasks round-robin with cp.async
        // input prefetch — faster alone under ncu, slower inside the back-to-back steps of bench.py.)
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(a.sort_ctl + 1, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= ntasks) break;
            // claim order: the LAST 32 entries of every chunk first, the first ones last — the entries of a chunk are in
            // ascending order of their bound, and for the exp families a larger bound means a deeper stop level, so
            // the kernel ends on its cheapest tasks (a warp that draws a task near the end delays the grid by half as much)
            const int slot = (SORT_CHUNK / 32 - 1) - t / nchunks, chunk = t - (t / nchunks) * nchunks;
            const int b = sorted[(long long)chunk * SORT_CHUNK + slot * 32 + lane];
            if (b >= 0) {
                T p[F::NP > 0 ? F::NP : 1];
                load_params<F::NP>(p, a, b);
                T dx, x0;
                bool neg;
                int flags;
                if (tpi_finish_prologue<T, F>(a, b, a.low[(long long)b * a.bstride], a.up[(long long)b * a.bstride], dx, x0, neg, flags)) {
                    if (F::template exp_arg_bound<T>(fm::abs(x0) + fm::abs(dx), p) < 700.0)
                        adaptive1d_tpi_levels<T, F, FAST, EXP_SMEM8 | EXP_SAFE>(a, b, p, dx, x0, neg, flags, a.tail_count, a.tail_idx);
                    else
                        adaptive1d_tpi_levels<T, F, FAST, EXP_SMEM8>(a, b, p, dx, x0, neg, flags, a.tail_count, a.tail_idx);
                }
            }
            __syncwarp();               // the claim is a full-warp shuffle
        }
    }
}

// Last kernel of a 1D thread-per-integral call (launched as a programmatic dependent of the previous one): the
// deep queue — integrals that were still unconverged at Args::tail_level — finished one CTA each: the CTA
// evaluates a tile of w_i*(f(x+)+f(x-)) terms in parallel into shared memory, then ONE thread adds them in index
// order, so the result is bit-identical to the sequential reference loop (adaptive.jl:55-60 / 763-770) while the
// integrand work is spread over 256 threads (Double64 tiles are tree-reduced instead, see tail_integral).
template <class T, class F, bool FAST, bool CMPL> __global__ void __launch_bounds__(256) k_adaptive1d_tail(const Args<T> a)
{
    family_tables_init<T, F>();
    constexpr int TILE = 1024;
    __shared__ T terms[TILE];
    __shared__ int s_done;
    const int tid = threadIdx.x;
    pdl_wait_primary();  // the kernels before this one have completed: the queue is final and visible
    const int count = *a.tail_count;
    for (int q = blockIdx.x; q < count; q += gridDim.x) tail_integral<T, F, FAST, CMPL>(a, a.tail_idx[q], terms, &s_done);
    // the last CTA to leave re-arms the (persistent, per-stream) queues for the next launch
    if (tid == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(a.ticket, 1u);
        if (t == gridDim.x - 1) {
            *a.tail_count = 0;
            *a.ticket = 0u;
            // ~ incoherent warps of this call: exact from k_bucket_sort_all, sampled (and never 0 when one was seen) from k_adaptive1d_tpi
            if (a.sort_fb) *a.sort_fb = a.sort_ctl[0] == 2 ? a.sort_ctl[2] : (a.sort_ctl[0] ? (a.sort_ctl[2] > 0 ? a.sort_ctl[2] : 1) : 0);
            if (a.sort_ctl) a.sort_ctl[0] = a.sort_ctl[1] = a.sort_ctl[2] = 0;
        }
    }
}

// integrand functors of kind FTS_KIND_2D_CMPL specialise this to true
template <class F> struct is_cmpl2d { static constexpr bool value = false; };

// ---- 2D helpers (adaptive.jl:165-176)
template <class T> struct Box2 { T dx, x0, dy, y0, w0; };

template <class T, class F, bool FAST> __device__ __forceinline__ T quadrants(const Box2<T>& bx, const T* p, T xi, T yi, T wi, T wj)
{
    const T xp = bx.dx * xi + bx.x0, xm = bx.x0 - bx.dx * xi;
    const T yp = bx.dy * yi + bx.y0, ym = bx.y0 - bx.dy * yi;
    return wi * wj * (F::template eval<T, FAST>(xp, yp, p) + F::template eval<T, FAST>(xm, yp, p) + F::template eval<T, FAST>(xp, ym, p) + F::template eval<T, FAST>(xm, ym, p));
}
template <class T, class F, bool FAST> __device__ __forceinline__ T axes2(const Box2<T>& bx, const T* p, T v, T wk)
{
    const T xp = bx.dx * v + bx.x0, xm = bx.x0 - bx.dx * v;
    const T yp = bx.dy * v + bx.y0, ym = bx.y0 - bx.dy * v;
    return wk * bx.w0 * (F::template eval<T, FAST>(xp, bx.y0, p) + F::template eval<T, FAST>(xm, bx.y0, p) + F::template eval<T, FAST>(bx.x0, yp, p) + F::template eval<T, FAST>(bx.x0, ym, p));
}
// tensor-complement twins (extension for BASELINE config C; coordinates as adaptive.jl:764-769 per axis)
template <class T, class F, bool FAST> __device__ __forceinline__ T quadrants_c(const Box2<T>& bx, const T* p, T xi, T ci, T yj, T cj, T wi, T wj)
{
    const T one = num<T>::one();
    const T xp = bx.dx * xi + bx.x0, xm = bx.x0 - bx.dx * xi;
    const T yp = bx.dy * yj + bx.y0, ym = bx.y0 - bx.dy * yj;
    const T bxp = bx.dx * ci, axp = bx.dx * (one + xi);
    const T dyp = bx.dy * cj, cyp = bx.dy * (one + yj);
    return wi * wj * (F::template eval<T, FAST>(xp, yp, bxp, axp, dyp, cyp, p) + F::template eval<T, FAST>(xm, yp, axp, bxp, dyp, cyp, p) +
                      F::template eval<T, FAST>(xp, ym, bxp, axp, cyp, dyp, p) + F::template eval<T, FAST>(xm, ym, axp, bxp, cyp, dyp, p));
}
template <class T, class F, bool FAST> __device__ __forceinline__ T axes2_c(const Box2<T>& bx, const T* p, T v, T c, T wk)
{
    const T one = num<T>::one();
    const T xp = bx.dx * v + bx.x0, xm = bx.x0 - bx.dx * v;
    const T yp = bx.dy * v + bx.y0, ym = bx.y0 - bx.dy * v;
    const T bxc = bx.dx * c, ax = bx.dx * (one + v);
    const T dyc = bx.dy * c, cy = bx.dy * (one + v);
    return wk * bx.w0 * (F::template eval<T, FAST>(xp, bx.y0, bxc, ax, bx.dy, bx.dy, p) + F::template eval<T, FAST>(xm, bx.y0, ax, bxc, bx.dy, bx.dy, p) +
                         F::template eval<T, FAST>(bx.x0, yp, bx.dx, bx.dx, dyc, cy, p) + F::template eval<T, FAST>(bx.x0, ym, bx.dx, bx.dx, cy, dyc, p));
}

// common prologue for 2D/3D boxes: loads bounds, applies quad edge rules (quad.jl:96-115,129-153)
template <int D, class T> __device__ __forceinline__ bool load_box(const Args<T>& a, long long b, T (&lo)[D], T (&hi)[D], bool& neg, int& flags)
{
#pragma unroll
    for (int d = 0; d < D; ++d) {
        lo[d] = a.low[b * a.bstride + d];
        hi[d] = a.up[b * a.bstride + d];
    }
    neg = false;
    flags = 0;
    if (a.options & OPT_QUAD_EDGE) {
        bool zero = false;
#pragma unroll
        for (int d = 0; d < D; ++d) zero = zero || (lo[d] == hi[d]);
        if (zero) return false;
#pragma unroll
        for (int d = 0; d < D; ++d) quad_edge_axis(lo[d], hi[d], neg);
        if (neg) flags |= FLAG_FLIPPED;
    }
    return true;
}

// adaptive_integrate_2D typed core   /root/reference/src/adaptive.jl:154-224
// CMPL=true: tensor-complement extension on NodeC tables (BASELINE config C)
template <class T, class F, bool FAST> __global__ void __launch_bounds__(128) k_adaptive2d_tpi(const Args<T> a)
{
    family_tables_init<T, F>();
    constexpr bool CMPL = is_cmpl2d<F>::value;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    T lo[2], hi[2];
    bool neg;
    int flags;
    if (!load_box<2>(a, b, lo, hi, neg, flags)) {
        store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
        return;
    }
    const T half = num<T>::half();
    Box2<T> bx;
    bx.dx = half * (hi[0] - lo[0]); bx.x0 = half * (hi[0] + lo[0]);
    bx.dy = half * (hi[1] - lo[1]); bx.y0 = half * (hi[1] + lo[1]);
    bx.w0 = num<T>::half_pi();
    T h = a.tm * half;
    const T w02 = bx.w0 * bx.w0;
    T s_total;
    if constexpr (CMPL) s_total = w02 * F::template eval<T, FAST>(bx.x0, bx.y0, bx.dx, bx.dx, bx.dy, bx.dy, p);
    else s_total = w02 * F::template eval<T, FAST>(bx.x0, bx.y0, p);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if constexpr (CMPL) s_total = s_total + quadrants_c<T, F, FAST>(bx, p, a.ix[i], a.ic[i], a.ix[j], a.ic[j], a.iw[i], a.iw[j]);
            else s_total = s_total + quadrants<T, F, FAST>(bx, p, a.ix[i], a.ix[j], a.iw[i], a.iw[j]);
        }
        if constexpr (CMPL) s_total = s_total + axes2_c<T, F, FAST>(bx, p, a.ix[i], a.ic[i], a.iw[i]);
        else s_total = s_total + axes2<T, F, FAST>(bx, p, a.ix[i], a.iw[i]);
    }
    T old_res = bx.dx * bx.dy * (h * h) * s_total;
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        T s_new = num<T>::zero();
        const int n = 2 << level;
        const int off = n - 4;
        for (int i = 1; i <= n; ++i) {
            T xi, wi, ci = num<T>::zero();
            if constexpr (CMPL) { const NodeC<T> nd = a.nodes_c[off + i - 1]; xi = nd.x; wi = nd.w; ci = nd.c; }
            else { const Node<T> nd = ldg_node(a.nodes + off + i - 1); xi = nd.x; wi = nd.w; }
            const int jstep = (i & 1) ? 1 : 2;  // even i: only odd j are new (adaptive.jl:203)
            for (int j = 1; j <= n; j += jstep) {
                if constexpr (CMPL) { const NodeC<T> nj = a.nodes_c[off + j - 1]; s_new = s_new + quadrants_c<T, F, FAST>(bx, p, xi, ci, nj.x, nj.c, wi, nj.w); }
                else { const Node<T> nj = ldg_node(a.nodes + off + j - 1); s_new = s_new + quadrants<T, F, FAST>(bx, p, xi, nj.x, wi, nj.w); }
            }
            if (i & 1) {
                if constexpr (CMPL) s_new = s_new + axes2_c<T, F, FAST>(bx, p, xi, ci, wi);
                else s_new = s_new + axes2<T, F, FAST>(bx, p, xi, wi);
            }
        }
        s_total = s_total + s_new;
        const T new_res = bx.dx * bx.dy * (h * h) * s_total;
        err = fm::abs(new_res - old_res);
        level_used = level;
        old_res = new_res;
        if (err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels)) {
            flags |= FLAG_CONVERGED;
            break;
        }
    }
    if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
    store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
}

// ---- 3D helpers (adaptive.jl:365-403)
template <class T> struct Box3 { T dx, x0, dy, y0, dz, z0, w0, w02; };

template <class T, class F, bool FAST> __device__ __forceinline__ T octant(const Box3<T>& bx, const T* p, T vi, T vj, T vk, T wi, T wj, T wk)
{
    const T xp = bx.dx * vi + bx.x0, xm = bx.x0 - bx.dx * vi;
    const T yp = bx.dy * vj + bx.y0, ym = bx.y0 - bx.dy * vj;
    const T zp = bx.dz * vk + bx.z0, zm = bx.z0 - bx.dz * vk;
    const T w = wi * wj * wk;
#define FE(X, Y, Z) F::template eval<T, FAST>(X, Y, Z, p)
    return w * ((FE(xp, yp, zp) + FE(xm, yp, zp) + FE(xp, ym, zp) + FE(xm, ym, zp)) + (FE(xp, yp, zm) + FE(xm, yp, zm) + FE(xp, ym, zm) + FE(xm, ym, zm)));
}
template <class T, class F, bool FAST> __device__ __forceinline__ T planes(const Box3<T>& bx, const T* p, T vi, T vj, T wi, T wj)
{
    const T x0 = bx.x0, y0 = bx.y0, z0 = bx.z0;
    const T xp = bx.dx * vi + x0, xm = x0 - bx.dx * vi;
    const T ypi = bx.dy * vi + y0, ymi = y0 - bx.dy * vi;
    const T ypj = bx.dy * vj + y0, ymj = y0 - bx.dy * vj;
    const T zpj = bx.dz * vj + z0, zmj = z0 - bx.dz * vj;
    const T w = wi * wj * bx.w0;
    return w * ((FE(xp, ypj, z0) + FE(xm, ypj, z0) + FE(xp, ymj, z0) + FE(xm, ymj, z0)) + (FE(xp, y0, zpj) + FE(xm, y0, zpj) + FE(xp, y0, zmj) + FE(xm, y0, zmj)) +
                (FE(x0, ypi, zpj) + FE(x0, ymi, zpj) + FE(x0, ypi, zmj) + FE(x0, ymi, zmj)));
}
template <class T, class F, bool FAST> __device__ __forceinline__ T axes3(const Box3<T>& bx, const T* p, T vi, T wi)
{
    const T x0 = bx.x0, y0 = bx.y0, z0 = bx.z0;
    const T xp = bx.dx * vi + x0, xm = x0 - bx.dx * vi;
    const T yp = bx.dy * vi + y0, ym = y0 - bx.dy * vi;
    const T zp = bx.dz * vi + z0, zm = z0 - bx.dz * vi;
    const T w = wi * bx.w02;
    return w * ((FE(xp, y0, z0) + FE(xm, y0, z0)) + (FE(x0, yp, z0) + FE(x0, ym, z0)) + (FE(x0, y0, zp) + FE(x0, y0, zm)));
}
#undef FE

template <class T> __device__ __forceinline__ Box3<T> make_box3(const T* lo, const T* hi)
{
    const T half = num<T>::half();
    Box3<T> bx;
    bx.dx = half * (hi[0] - lo[0]); bx.x0 = half * (hi[0] + lo[0]);
    bx.dy = half * (hi[1] - lo[1]); bx.y0 = half * (hi[1] + lo[1]);
    bx.dz = half * (hi[2] - lo[2]); bx.z0 = half * (hi[2] + lo[2]);
    bx.w0 = num<T>::half_pi();
    bx.w02 = bx.w0 * bx.w0;
    return bx;
}

// level-0 sum (adaptive.jl:406-418)
template <class T, class F, bool FAST> __device__ __forceinline__ T level0_3d(const Box3<T>& bx, const T* p, const T (&ix)[2], const T (&iw)[2])
{
    const T w03 = bx.w02 * bx.w0;
    T s_total = w03 * F::template eval<T, FAST>(bx.x0, bx.y0, bx.z0, p);
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
#pragma unroll 1
            for (int k = 0; k < 2; ++k) s_total = s_total + octant<T, F, FAST>(bx, p, ix[i], ix[j], ix[k], iw[i], iw[j], iw[k]);
            s_total = s_total + planes<T, F, FAST>(bx, p, ix[i], ix[j], iw[i], iw[j]);
        }
        s_total = s_total + axes3<T, F, FAST>(bx, p, ix[i], iw[i]);
    }
    return s_total;
}

// adaptive_integrate_3D typed core   /root/reference/src/adaptive.jl:352-471
template <class T, class F, bool FAST> __global__ void __launch_bounds__(128) k_adaptive3d_tpi(const Args<T> a)
{
    family_tables_init<T, F>();
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    T lo[3], hi[3];
    bool neg;
    int flags;
    if (!load_box<3>(a, b, lo, hi, neg, flags)) {
        store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
        return;
    }
    const T half = num<T>::half();
    const Box3<T> bx = make_box3(lo, hi);
    T h = a.tm * half;
    T s_total = level0_3d<T, F, FAST>(bx, p, a.ix, a.iw);
    T old_res = bx.dx * bx.dy * bx.dz * (h * h * h) * s_total;
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        T s_new = num<T>::zero();
        const int n = 2 << level;
        const Node<T>* __restrict__ lv = a.nodes + (n - 4);
        for (int i = 1; i <= n; ++i) {
            const Node<T> ni = ldg_node(lv + i - 1);
            for (int j = 1; j <= n; ++j) {
                const Node<T> nj = ldg_node(lv + j - 1);
                const bool full = (i & 1) || (j & 1);  // adaptive.jl:432-453 index classes
                const int kstep = full ? 1 : 2;
#pragma unroll 1
                for (int k = 1; k <= n; k += kstep) {
                    const Node<T> nk = ldg_node(lv + k - 1);
                    s_new = s_new + octant<T, F, FAST>(bx, p, ni.x, nj.x, nk.x, ni.w, nj.w, nk.w);
                }
                if (full) s_new = s_new + planes<T, F, FAST>(bx, p, ni.x, nj.x, ni.w, nj.w);
            }
            if (i & 1) s_new = s_new + axes3<T, F, FAST>(bx, p, ni.x, ni.w);
        }
        s_total = s_total + s_new;
        const T new_res = bx.dx * bx.dy * bx.dz * (h * h * h) * s_total;
        err = fm::abs(new_res - old_res);
        level_used = level;
        old_res = new_res;
        if (err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels)) {
            flags |= FLAG_CONVERGED;
            break;
        }
    }
    if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
    store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
}

// ============================================== reference order at cooperative speed (2D / 3D) ====
// The scalar 2D/3D kernels of the reference add every term of a level to ONE accumulator in loop
// order (adaptive.jl:199-209, 430-456), so the sum cannot be split without changing its rounding.
// What can be split is the evaluation: a CTA owns one integral, its 256 threads evaluate a tile of
// the level's term SEQUENCE (cell, plane and axis terms at their positions in the loop nest, each by
// the same device function the thread-per-integral kernel calls), and one thread adds the tile in
// order.  Bit-identical to k_adaptive{2,3}d_tpi<.., false>; the dependent add (8 cycles per term)
// replaces a dependent chain of whole integrand evaluations, ~20-50x faster for one integral.

// ---- thread-block cluster plumbing of the ordered adaptive kernels ------------------------------
// One integral may be served by a CLUSTER of up to 8 CTAs: every CTA evaluates terms and stores them
// through distributed shared memory into the tile buffers of the cluster's CTA 0, whose thread 0 is
// the only one that adds — while it adds tile t, everyone else already evaluates tile t+1.  The
// ordered add (8 cycles per term) then is the only exposed cost: one 3D integral at level 6,
// 40.6 -> 9 ms.  With a cluster of 1 (the default launch) the same code runs inside one CTA.
extern __shared__ __align__(16) unsigned char fts_dyn_smem[];
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync()
{
    if (cluster_nctarank() == 1) {  // a lone CTA: the CTA barrier is cheaper than the cluster barrier
        __syncthreads();
        return;
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory pointer of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ unsigned cluster_map(const void* p, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_store(unsigned addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void cluster_store(unsigned addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void cluster_store_int(unsigned addr, int v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// Distributed-shared-memory lifetime rules the ordered kernels keep:
//  (1) no CTA touches a peer's shared memory before cluster_entry_sync() — until then the peer may
//      not have started;
//  (2) nobody READS remote shared memory at all: the adder pushes the stop flag into every CTA's own
//      s_done before the level's closing barrier, so after the last barrier a CTA only looks at its
//      own memory and CTA 0 may retire while its peers are still leaving.
__device__ __forceinline__ void cluster_entry_sync()
{
    if (cluster_nctarank() > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// adder thread: publish `v` to the s_done word of every CTA of the cluster (own CTA included)
__device__ __forceinline__ void cluster_broadcast_int(int* local_word, int v)
{
    const unsigned csize = cluster_nctarank();
    for (unsigned r = 0; r < csize; ++r) cluster_store_int(cluster_map(local_word, r), v);
}

// Ordered sum of term(0), term(1), ..., term(nterms-1) into s (valid in the adder thread: thread 0 of
// cluster CTA 0).  Every thread of every CTA of the cluster calls this with the same arguments.
template <class T, class TermFn> __device__ __forceinline__ void ordered_sum_cluster(T& s, long long nterms, int tile_terms, TermFn term)
{
    const int ORDC_TILE = tile_terms;  // terms per buffer (Args::level): small for many integrals (more CTAs per SM), large for a few
    T* bufs = reinterpret_cast<T*>(fts_dyn_smem);  // [2][tile_terms] in dynamic shared memory, used in CTA 0 only
    const unsigned crank = cluster_ctarank(), csize = cluster_nctarank();
    const unsigned remote = cluster_map(bufs, 0);
    const bool adder = crank == 0 && threadIdx.x == 0;
    const int wid = (int)(crank * blockDim.x + threadIdx.x) - 1;  // evaluator index; the adder does not evaluate
    const int nw = (int)(csize * blockDim.x) - 1;
    const long long ntile = (nterms + ORDC_TILE - 1) / ORDC_TILE;
    auto evaluate = [&](long long t) {
        const long long t0 = t * ORDC_TILE;
        const int m = (int)(nterms - t0 < ORDC_TILE ? nterms - t0 : ORDC_TILE);
        const unsigned base = remote + (unsigned)((t & 1) * ORDC_TILE * sizeof(T));
        for (int i = wid; i < m; i += nw) cluster_store(base + (unsigned)(i * sizeof(T)), term(t0 + i));
    };
    if (!adder && ntile > 0) evaluate(0);
    cluster_sync();
    for (long long t = 0; t < ntile; ++t) {
        if (adder) {
            const long long t0 = t * ORDC_TILE;
            const int m = (int)(nterms - t0 < ORDC_TILE ? nterms - t0 : ORDC_TILE);
            ordered_add(s, bufs + (t & 1) * ORDC_TILE, m);
        } else if (t + 1 < ntile) {
            evaluate(t + 1);
        }
        cluster_sync();
    }
}


// term `pos` of level-l sequence of the 2D loop nest: for i = 1..n { cells (i, j), j = 1..n (all j
// if i is odd, odd j otherwise); axis term of i if i is odd }
template <class T, class F, bool CMPL> __device__ __forceinline__ T term2d(const Args<T>& a, const Box2<T>& bx, const T* p, int n, long long pos)
{
    const int off = n - 4;
    const int P = n + 1 + (n >> 1);  // an odd row (n cells + axis) followed by an even row (n/2 cells)
    const int q = (int)(pos / P), r = (int)(pos - (long long)q * P);
    int i, j;
    bool axis = false;
    if (r <= n) {
        i = 2 * q + 1;
        j = r + 1;
        axis = r == n;
    } else {
        i = 2 * q + 2;
        j = 2 * (r - (n + 1)) + 1;
    }
    if constexpr (CMPL) {
        const NodeC<T> ni = a.nodes_c[off + i - 1];
        if (axis) return axes2_c<T, F, false>(bx, p, ni.x, ni.c, ni.w);
        const NodeC<T> nj = a.nodes_c[off + j - 1];
        return quadrants_c<T, F, false>(bx, p, ni.x, ni.c, nj.x, nj.c, ni.w, nj.w);
    } else {
        const Node<T> ni = ldg_node(a.nodes + off + i - 1);
        if (axis) return axes2<T, F, false>(bx, p, ni.x, ni.w);
        const Node<T> nj = ldg_node(a.nodes + off + j - 1);
        return quadrants<T, F, false>(bx, p, ni.x, nj.x, ni.w, nj.w);
    }
}

// adaptive_integrate_2D typed core, reference order, one CTA per integral   adaptive.jl:154-224
template <class T, class F> __global__ void __launch_bounds__(256) k_adaptive2d_ord(const Args<T> a)
{
    family_tables_init<T, F>();
    constexpr bool CMPL = is_cmpl2d<F>::value;
    __shared__ int s_done;
    const long long b = cluster_id_x();  // one integral per cluster (of 1, 2, 4 or 8 CTAs)
    const bool adder = cluster_ctarank() == 0 && threadIdx.x == 0;
    const int tid = adder ? 0 : 1;  // "thread 0" below = the adder thread of the cluster
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    T lo[2], hi[2];
    bool neg;
    int flags;
    if (!load_box<2>(a, b, lo, hi, neg, flags)) {
        if (tid == 0) store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
        return;  // the whole cluster leaves (same box in every CTA), nothing remote was touched
    }
    cluster_entry_sync();
    const T half = num<T>::half();
    Box2<T> bx;
    bx.dx = half * (hi[0] - lo[0]); bx.x0 = half * (hi[0] + lo[0]);
    bx.dy = half * (hi[1] - lo[1]); bx.y0 = half * (hi[1] + lo[1]);
    bx.w0 = num<T>::half_pi();
    T h = a.tm * half;
    T s_total = num<T>::zero();
    if (tid == 0) {  // level 0 exactly as k_adaptive2d_tpi (25 evaluations)
        const T w02 = bx.w0 * bx.w0;
        if constexpr (CMPL) s_total = w02 * F::template eval<T, false>(bx.x0, bx.y0, bx.dx, bx.dx, bx.dy, bx.dy, p);
        else s_total = w02 * F::template eval<T, false>(bx.x0, bx.y0, p);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if constexpr (CMPL) s_total = s_total + quadrants_c<T, F, false>(bx, p, a.ix[i], a.ic[i], a.ix[j], a.ic[j], a.iw[i], a.iw[j]);
                else s_total = s_total + quadrants<T, F, false>(bx, p, a.ix[i], a.ix[j], a.iw[i], a.iw[j]);
            }
            if constexpr (CMPL) s_total = s_total + axes2_c<T, F, false>(bx, p, a.ix[i], a.ic[i], a.iw[i]);
            else s_total = s_total + axes2<T, F, false>(bx, p, a.ix[i], a.iw[i]);
        }
    }
    T old_res = bx.dx * bx.dy * (h * h) * s_total;
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        const int n = 2 << level;
        const long long nterms = (long long)(n >> 1) * (n + 1 + (n >> 1));
        T s_new = num<T>::zero();
        ordered_sum_cluster(s_new, nterms, a.level, [&](long long pos) { return term2d<T, F, CMPL>(a, bx, p, n, pos); });
        if (tid == 0) {
            s_total = s_total + s_new;
            const T new_res = bx.dx * bx.dy * (h * h) * s_total;
            err = fm::abs(new_res - old_res);
            level_used = level;
            old_res = new_res;
            const bool conv = err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels);
            if (conv) flags |= FLAG_CONVERGED;
            cluster_broadcast_int(&s_done, conv ? 1 : 0);
        }
        cluster_sync();
        if (*(volatile int*)&s_done) break;
    }
    if (tid == 0) {
        if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
}

// term `pos` of the 3D loop nest: for i { for j { cells (i,j,k): all k if i or j is odd, odd k
// otherwise; plane term of (i,j) if i or j is odd }; axis term of i if i is odd }
template <class T, class F> __device__ __forceinline__ T term3d(const Args<T>& a, const Box3<T>& bx, const T* p, int n, long long pos)
{
    const Node<T>* __restrict__ lv = a.nodes + (n - 4);
    const long long b_odd = (long long)n * (n + 1) + 1;          // odd i: n full rows + the axis term
    const int R = (n + 1) + (n >> 1);                            // even i: a full row (odd j) + a half row (even j)
    const long long b_even = (long long)(n >> 1) * R;
    const long long B = b_odd + b_even;
    const long long qi = pos / B;
    long long r = pos - qi * B;
    int i, j, k;
    bool plane = false;
    if (r < b_odd) {
        i = (int)(2 * qi + 1);
        if (r == b_odd - 1) {
            const Node<T> ni = ldg_node(lv + i - 1);
            return axes3<T, F, false>(bx, p, ni.x, ni.w);
        }
        j = (int)(r / (n + 1)) + 1;
        const int rr = (int)(r - (long long)(j - 1) * (n + 1));
        plane = rr == n;
        k = rr + 1;
    } else {
        r -= b_odd;
        i = (int)(2 * qi + 2);
        const int qj = (int)(r / R), r2 = (int)(r - (long long)qj * R);
        if (r2 <= n) {
            j = 2 * qj + 1;
            plane = r2 == n;
            k = r2 + 1;
        } else {
            j = 2 * qj + 2;
            k = 2 * (r2 - (n + 1)) + 1;
        }
    }
    const Node<T> ni = ldg_node(lv + i - 1), nj = ldg_node(lv + j - 1);
    if (plane) return planes<T, F, false>(bx, p, ni.x, nj.x, ni.w, nj.w);
    const Node<T> nk = ldg_node(lv + k - 1);
    return octant<T, F, false>(bx, p, ni.x, nj.x, nk.x, ni.w, nj.w, nk.w);
}

// adaptive_integrate_3D typed core, reference order, one CTA per integral   adaptive.jl:352-471
template <class T, class F> __global__ void __launch_bounds__(256) k_adaptive3d_ord(const Args<T> a)
{
    family_tables_init<T, F>();
    __shared__ int s_done;
    const long long b = cluster_id_x();  // one integral per cluster (of 1, 2, 4 or 8 CTAs)
    const bool adder = cluster_ctarank() == 0 && threadIdx.x == 0;
    const int tid = adder ? 0 : 1;  // "thread 0" below = the adder thread of the cluster
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    T lo[3], hi[3];
    bool neg;
    int flags;
    if (!load_box<3>(a, b, lo, hi, neg, flags)) {
        if (tid == 0) store_result(a, b, num<T>::zero(), num<T>::zero(), 0, FLAG_ZERO_WIDTH | FLAG_CONVERGED);
        return;  // the whole cluster leaves (same box in every CTA), nothing remote was touched
    }
    cluster_entry_sync();
    const T half = num<T>::half();
    const Box3<T> bx = make_box3(lo, hi);
    T h = a.tm * half;
    T s_total = num<T>::zero();
    if (tid == 0) s_total = level0_3d<T, F, false>(bx, p, a.ix, a.iw);
    T old_res = bx.dx * bx.dy * bx.dz * (h * h * h) * s_total;
    T err = num<T>::zero();
    int level_used = 0;
    for (int level = 1; level <= a.max_levels; ++level) {
        h = h * half;
        const int n = 2 << level;
        const long long nterms = (long long)(n >> 1) * ((long long)n * (n + 1) + 1 + (long long)(n >> 1) * ((n + 1) + (n >> 1)));
        T s_new = num<T>::zero();
        ordered_sum_cluster(s_new, nterms, a.level, [&](long long pos) { return term3d<T, F>(a, bx, p, n, pos); });
        if (tid == 0) {
            s_total = s_total + s_new;
            const T new_res = bx.dx * bx.dy * bx.dz * (h * h * h) * s_total;
            err = fm::abs(new_res - old_res);
            level_used = level;
            old_res = new_res;
            const bool conv = err <= fm::err_target(new_res, a.rtol, a.atol) && may_stop(a.options, level, a.max_levels);
            if (conv) flags |= FLAG_CONVERGED;
            cluster_broadcast_int(&s_done, conv ? 1 : 0);
        }
        cluster_sync();
        if (*(volatile int*)&s_done) break;
    }
    if (tid == 0) {
        if (!(flags & FLAG_CONVERGED) && a.max_levels > 0) flags |= FLAG_MAX_LEVELS;
        store_result(a, b, neg ? -old_res : old_res, err, level_used, flags);
    }
}

// ---- fixed grids in reference order, one CTA per integral --------------------------------------
// integrate2D / integrate3D (integrate2D.jl:91-124, integrate3D.jl:150-207) sum row by row: an inner
// sum over j (2D) or k (3D) per row, rows folded into the total in loop order.  Rows are independent,
// so every thread walks whole rows with the inner loop of k_fixed{2,3}d_tpi, the row results go to
// shared memory, and one thread folds them in order: bit-identical, and the dependent chain is one
// row instead of the whole grid (one integrate3D at N = 256: 17 M evaluations, 0.35 s on one thread).

// integrate1D in reference order (integrate1D.jl:90-102): terms in parallel, added in order
// (cluster machinery of ordered_sum_cluster: one cluster of 1..8 CTAs per integral)
template <class T, class F> __global__ void __launch_bounds__(256) k_fixed1d_ord(const Args<T> a)
{
    family_tables_init<T, F>();
    cluster_entry_sync();
    const long long b = cluster_id_x();
    const bool adder = cluster_ctarank() == 0 && threadIdx.x == 0;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T lo = a.low[b * a.bstride], hi = a.up[b * a.bstride];
    const T dx = half * (hi - lo), x0 = half * (hi + lo);
    T s = num<T>::zero();
    if (adder) s = num<T>::half_pi() * F::template eval<T, false>(x0, p);
    ordered_sum_cluster(s, (long long)a.n, a.level, [&](long long pos) {
        const Node<T> nd = ldg_node(a.grid + pos);
        const T d = dx * nd.x;
        const T xm = x0 - d, xp = x0 + d;
        return nd.w * (F::template eval<T, false>(xm, p) + F::template eval<T, false>(xp, p));
    });
    if (adder) a.value[b] = dx * a.h * s;
}

// integrate2D: n axis terms, then n row terms w_i * (inner sum over j)
template <class T, class F> __global__ void __launch_bounds__(256) k_fixed2d_ord(const Args<T> a)
{
    family_tables_init<T, F>();
    cluster_entry_sync();
    const long long b = cluster_id_x();
    const bool adder = cluster_ctarank() == 0 && threadIdx.x == 0;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T* lo = a.low + b * a.bstride;
    const T* hi = a.up + b * a.bstride;
    const T dx = half * (hi[0] - lo[0]), x0 = half * (hi[0] + lo[0]);
    const T dy = half * (hi[1] - lo[1]), y0 = half * (hi[1] + lo[1]);
    const T w0 = num<T>::half_pi();
    const T w02 = w0 * w0;
    const int n = a.n;
    T s = num<T>::zero();
    if (adder) s = w02 * F::template eval<T, false>(x0, y0, p);
    ordered_sum_cluster(s, 2LL * n, a.level, [&](long long pos) {
        if (pos < n) {
            const Node<T> nd = ldg_node(a.grid + pos);
            const T xkp = dx * nd.x + x0, xkm = x0 - dx * nd.x;
            const T ykp = dy * nd.x + y0, ykm = y0 - dy * nd.x;
            return nd.w * w0 * (F::template eval<T, false>(xkp, y0, p) + F::template eval<T, false>(xkm, y0, p) +
                                F::template eval<T, false>(x0, ykp, p) + F::template eval<T, false>(x0, ykm, p));
        }
        const Node<T> ni = ldg_node(a.grid + (pos - n));
        const T xip = dx * ni.x + x0, xim = x0 - dx * ni.x;
        T inner = num<T>::zero();
        for (int j = 0; j < n; ++j) {
            const Node<T> nj = ldg_node(a.grid + j);
            const T yjp = dy * nj.x + y0, yjm = y0 - dy * nj.x;
            inner = inner + nj.w * (F::template eval<T, false>(xim, yjm, p) + F::template eval<T, false>(xip, yjm, p) +
                                    F::template eval<T, false>(xim, yjp, p) + F::template eval<T, false>(xip, yjp, p));
        }
        return ni.w * inner;
    });
    if (adder) a.value[b] = dx * dy * (a.h * a.h) * s;
}

// integrate3D: n axis terms, then for every (i, j) its plane term followed by its octant row sum
template <class T, class F> __global__ void __launch_bounds__(256) k_fixed3d_ord(const Args<T> a)
{
    family_tables_init<T, F>();
    cluster_entry_sync();
    const long long b = cluster_id_x();
    const bool adder = cluster_ctarank() == 0 && threadIdx.x == 0;
    T p[F::NP > 0 ? F::NP : 1];
    load_params<F::NP>(p, a, b);
    const T half = num<T>::half();
    const T* lo = a.low + b * a.bstride;
    const T* hi = a.up + b * a.bstride;
    const T dx = half * (hi[0] - lo[0]), x0 = half * (hi[0] + lo[0]);
    const T dy = half * (hi[1] - lo[1]), y0 = half * (hi[1] + lo[1]);
    const T dz = half * (hi[2] - lo[2]), z0 = half * (hi[2] + lo[2]);
    const T w0 = num<T>::half_pi();
    const T w02 = w0 * w0;
    const T w03 = w02 * w0;
    const int n = a.n;
#define FE(X, Y, Z) F::template eval<T, false>(X, Y, Z, p)
    T total = num<T>::zero();
    if (adder) total = w03 * FE(x0, y0, z0);
    ordered_sum_cluster(total, (long long)n + 2LL * n * n, a.level, [&](long long pos) {
        if (pos < n) {
            const Node<T> nd = ldg_node(a.grid + pos);
            const T axis = (FE(dx * nd.x + x0, y0, z0) + FE(x0 - dx * nd.x, y0, z0)) + (FE(x0, dy * nd.x + y0, z0) + FE(x0, y0 - dy * nd.x, z0)) +
                           (FE(x0, y0, dz * nd.x + z0) + FE(x0, y0, z0 - dz * nd.x));
            return nd.w * w02 * axis;
        }
        const long long q = (pos - n) >> 1;
        const int i = (int)(q / n), j = (int)(q - (long long)i * n);
        const Node<T> ni = ldg_node(a.grid + i);
        const T xp = dx * ni.x + x0, xm = x0 - dx * ni.x;
        const Node<T> nj = ldg_node(a.grid + j);
        const T wiwj = ni.w * nj.w;
        const T yp = dy * nj.x + y0, ym = y0 - dy * nj.x;
        if (((pos - n) & 1) == 0) {
            const T ypi = dy * ni.x + y0, ymi = y0 - dy * ni.x;
            const T zpj = dz * nj.x + z0, zmj = z0 - dz * nj.x;
            const T plane = (FE(xp, yp, z0) + FE(xm, yp, z0) + FE(xp, ym, z0) + FE(xm, ym, z0)) +
                            (FE(xp, y0, zpj) + FE(xm, y0, zpj) + FE(xp, y0, zmj) + FE(xm, y0, zmj)) +
                            (FE(x0, ypi, zpj) + FE(x0, ymi, zpj) + FE(x0, ypi, zmj) + FE(x0, ymi, zmj));
            return wiwj * w0 * plane;
        }
        T oct = num<T>::zero();
        for (int k = 0; k < n; ++k) {
            const Node<T> nk = ldg_node(a.grid + k);
            const T zp = dz * nk.x + z0, zm = z0 - dz * nk.x;
            oct = oct + nk.w * ((FE(xp, yp, zp) + FE(xm, yp, zp) + FE(xp, ym, zp) + FE(xm, ym, zp)) +
                                (FE(xp, yp, zm) + FE(xm, yp, zm) + FE(xp, ym, zm) + FE(xm, ym, zm)));
        }
        return wiwj * oct;
    });
#undef FE
    if (adder) a.value[b] = (dx * dy * dz * (a.h * a.h * a.h)) * total;
}

}  // namespace fts
