// fts_host.cu — the C ABI of libfts_b200.so (include/fts_b200.h): lifecycle, table/cache handles,
// kernel registry and launch dispatch.  There is NO CPU fallback: every compute entry point
// requires an sm_100 device and fails with FTS_ERR_NO_DEVICE / FTS_ERR_CUDA otherwise.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "fts_internal.h"
#include "fts_kernels_cta.cuh"
#include "fts_tablegen_dev.cuh"

using namespace fts_host;

// tablegen (fts_tablegen.cpp)
extern "C" {
int fts_tg_window(int dtype, int which, int D, void* out);
int fts_tg_fixed(int dtype, int N, int D, void* x, void* w, void* h);
int fts_tg_cache1d(int dtype, int L, int complement, void* tm, void* ix, void* iw, void* ic, void* xs, void* ws, void* cs);
int fts_tg_cache_tensor(int dtype, int D, int L, void* tm, void* ix, void* iw, void* xs, void* ws);
int fts_tg_tensor_complements(int dtype, int L, const void* tm_in, void* ix, void* iw, void* ic, void* xs, void* ws, void* cs);
int fts_tg_hopt(int dtype, int N, double d, void* out);
int fts_tg_hmax(int dtype, int N, int D, void* out);
long fts_tg_opt(int dtype, int N, double d, int D, long capacity, void* x, void* w, void* h);
}

// ------------------------------------------------------------------------------- errors ------
namespace {
thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
std::mutex g_mu;  // lifecycle and cold paths only; the launch entry points do not take it
bool g_inited = false;
int g_device = -1;
cudaDeviceProp g_prop;
cudaStream_t g_stream = nullptr;  // stream of the cold paths (table generation, peak measurement), used under g_mu
unsigned long long* g_peer_wait = nullptr;  // [64] device: per-level exchange wait of the last fused sharded call (ns)

std::vector<KernelEntry>& registry()
{
    static std::vector<KernelEntry> r;
    return r;
}

const FamilyInfo k_families[] = {
    {FTS_F1_POLY7, FTS_KIND_1D, 8, "poly7"},
    {FTS_F1_EXP, FTS_KIND_1D, 0, "exp"},
    {FTS_F1_GAUSS, FTS_KIND_1D, 1, "gauss"},
    {FTS_F1_RUNGE, FTS_KIND_1D, 1, "runge"},
    {FTS_F1_LOG1MX, FTS_KIND_1D, 0, "log1mx"},
    {FTS_F1_LOG1PX, FTS_KIND_1D, 0, "log1px"},
    {FTS_F1_INVSQRT1MX2, FTS_KIND_1D, 0, "invsqrt1mx2"},
    {FTS_F1_SIN2, FTS_KIND_1D, 1, "sin2"},
    {FTS_F1_POLY6, FTS_KIND_1D, 0, "poly6"},
    {FTS_F1_INVSQRTABS, FTS_KIND_1D, 1, "invsqrtabs"},
    {FTS_C1_INVSQRT, FTS_KIND_1D_CMPL, 1, "c_invsqrt"},
    {FTS_C1_LOG_BMX, FTS_KIND_1D_CMPL, 0, "c_log_bmx"},
    {FTS_C1_LOG_XMA, FTS_KIND_1D_CMPL, 0, "c_log_xma"},
    {FTS_C1_EXP_INV_BMX, FTS_KIND_1D_CMPL, 0, "c_exp_inv_bmx"},
    {FTS_C1_EXP_PLUS, FTS_KIND_1D_CMPL, 0, "c_exp_plus"},
    {FTS_F2_POLY2, FTS_KIND_2D, 6, "poly2_2d"},
    {FTS_F2_EXP, FTS_KIND_2D, 0, "exp_2d"},
    {FTS_F2_X2PY2, FTS_KIND_2D, 0, "x2py2"},
    {FTS_F2_INVSQRT, FTS_KIND_2D, 0, "invsqrt_2d"},
    {FTS_F2_INVSQRTABS, FTS_KIND_2D, 0, "invsqrtabs_2d"},
    {FTS_C2_INVSQRT_BMX_YMC, FTS_KIND_2D_CMPL, 0, "c2_invsqrt_bmx_ymc"},
    {FTS_C2_INVSQRT_ALL, FTS_KIND_2D_CMPL, 0, "c2_invsqrt_all"},
    {FTS_F3_POLY2, FTS_KIND_3D, 6, "poly2_3d"},
    {FTS_F3_EXP, FTS_KIND_3D, 0, "exp_3d"},
    {FTS_F3_X2Y2Z2, FTS_KIND_3D, 0, "x2y2z2"},
    {FTS_F3_INVSQRT, FTS_KIND_3D, 0, "invsqrt_3d"},
    {FTS_F3_RAT, FTS_KIND_3D, 0, "rat_3d"},
    {FTS_F3_LOG, FTS_KIND_3D, 0, "log_3d"},
    {FTS_F3_INVSQRTABS, FTS_KIND_3D, 0, "invsqrtabs_3d"},
};

int kind_dim(int kind)
{
    switch (kind) {
    case FTS_KIND_1D: case FTS_KIND_1D_CMPL: return 1;
    case FTS_KIND_2D: case FTS_KIND_2D_CMPL: return 2;
    case FTS_KIND_3D: return 3;
    }
    if (kind > FTS_KIND_ND && kind <= FTS_KIND_ND + FTS_ND_MAX) return kind - FTS_KIND_ND;
    return 0;
}
bool kind_cmpl(int kind) { return kind == FTS_KIND_1D_CMPL || kind == FTS_KIND_2D_CMPL; }

// grow-only device scratch slots for the host-buffer entry points
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
};
enum { S_LOW, S_UP, S_PAR, S_VAL, S_ERR, S_LVL, S_FLG, S_COUNT };

// Host-buffer entry points (fts_adaptive_batch, fts_integrate_batch, ...) are re-entrant: every HOST
// THREAD owns a context — its own non-blocking stream, chunk-pipeline streams and grow-only device
// scratch — so two threads drive two streams side by side without a lock (the reference's kernels
// are lock-free too; only its cache memo takes a ReentrantLock, caches.jl:51).  Contexts are created
// on a thread's first call and re-armed after fts_shutdown / re-binding (generation counter).
struct HostCtx {
    unsigned long long gen = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev = nullptr;
    Scratch scratch[S_COUNT];
};
std::mutex g_ctx_mu;
std::vector<HostCtx*> g_ctxs;  // never freed: a thread keeps its pointer; the CUDA resources inside are dropped at shutdown
std::atomic<unsigned long long> g_gen{1};

void drop_host_contexts()  // fts_shutdown / re-binding; must not race with running calls
{
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (HostCtx* c : g_ctxs) {
        for (auto& sc : c->scratch) { if (sc.p) cudaFree(sc.p); sc.p = nullptr; sc.cap = 0; }
        if (c->stream) cudaStreamDestroy(c->stream);
        for (auto& ps : c->pipe) { if (ps) cudaStreamDestroy(ps); ps = nullptr; }
        if (c->ev) cudaEventDestroy(c->ev);
        c->stream = nullptr; c->ev = nullptr; c->gen = 0;
    }
    g_gen.fetch_add(1);
    cudaGetLastError();
}

int host_ctx(HostCtx** out)
{
    thread_local HostCtx* ctx = nullptr;
    if (!ctx) {
        ctx = new HostCtx();
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        g_ctxs.push_back(ctx);
    }
    if (ctx->gen != g_gen.load()) {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        FTS_CUDA(cudaSetDevice(g_device));  // the current device is a per-thread setting
        FTS_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        for (auto& ps : ctx->pipe) FTS_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
        FTS_CUDA(cudaEventCreateWithFlags(&ctx->ev, cudaEventDisableTiming));
        ctx->gen = g_gen.load();
    }
    *out = ctx;
    return FTS_OK;
}
}  // namespace

namespace fts_host {
int set_error(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}
int cuda_error(cudaError_t e, const char* what)
{
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? FTS_ERR_NO_DEVICE : FTS_ERR_CUDA;
}
void register_kernels(const KernelEntry* entries, int count)
{
    auto& r = registry();
    r.insert(r.end(), entries, entries + count);
}
const void* find_kernel(int family, int dtype, int kid, int fast)
{
    for (const auto& e : registry())
        if (e.family == family && e.dtype == dtype && e.kid == kid && e.fast == fast) return e.fn;
    return nullptr;
}
const FamilyInfo* family_info(int family)
{
    for (const auto& f : k_families)
        if (f.family == family) return &f;
    return nullptr;
}
}  // namespace fts_host

static void drop_device_state();  // per-stream queues and other device buffers owned by the library

static int ensure_init()
{
    if (g_inited) return FTS_OK;
    return fts_init(-1);
}

// ------------------------------------------------------------------------------- lifecycle ---
extern "C" int fts_abi_version(void) { return FTS_ABI_VERSION; }

extern "C" size_t fts_last_error(char* buf, size_t len)
{
    if (buf && len) {
        size_t n = g_last_error.size() < len - 1 ? g_last_error.size() : len - 1;
        memcpy(buf, g_last_error.data(), n);
        buf[n] = 0;
    }
    return g_last_error.size();
}

extern "C" int fts_init(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(FTS_ERR_NO_DEVICE,
                         "fts_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback");
    }
    if (device < 0) {
        if (g_inited) return FTS_OK;
        FTS_CUDA(cudaGetDevice(&device));
    }
    if (device >= count) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_init: device index out of range");
    FTS_CUDA(cudaSetDevice(device));
    FTS_CUDA(cudaGetDeviceProperties(&g_prop, device));
    if (g_prop.major != 10)
        return set_error(FTS_ERR_NO_DEVICE, std::string("fts_b200 kernels are built for sm_100a only; device is ") + g_prop.name);
    if (g_inited && g_device != device) {
        // re-binding: drop everything that lives on the old device
        drop_host_contexts();
        drop_device_state();
        if (g_stream) cudaStreamDestroy(g_stream);
        g_stream = nullptr;
    }
    if (!g_stream) {
        // stream-ordered scratch (cudaMallocAsync) must stay cached in the pool between calls:
        // with the default release threshold (0) every synchronisation returns it to the OS and
        // the next launch pays milliseconds to get it back
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        FTS_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    }
    g_device = device;
    g_inited = true;
    return FTS_OK;
}

extern "C" int fts_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) return FTS_OK;
    drop_host_contexts();
    drop_device_state();
    if (g_stream) cudaStreamDestroy(g_stream);
    g_stream = nullptr;
    g_inited = false;
    return FTS_OK;
}

extern "C" int fts_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor, size_t* global_mem)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (device) *device = g_device;
    if (sm_count) *sm_count = g_prop.multiProcessorCount;
    if (cc_major) *cc_major = g_prop.major;
    if (cc_minor) *cc_minor = g_prop.minor;
    if (global_mem) *global_mem = g_prop.totalGlobalMem;
    return FTS_OK;
}

extern "C" int fts_host_alloc(void** ptr, size_t bytes)
{
    int rc = ensure_init();
    if (rc) return rc;
    FTS_CUDA(cudaMallocHost(ptr, bytes));
    return FTS_OK;
}
extern "C" int fts_host_free(void* ptr)
{
    FTS_CUDA(cudaFreeHost(ptr));
    return FTS_OK;
}

extern "C" int64_t fts_launch_count(int reset)
{
    long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}

extern "C" int fts_family_info(int family, int* kind, int* nparams)
{
    const FamilyInfo* f = family_info(family);
    if (!f) return set_error(FTS_ERR_INVALID_ARGUMENT, "unknown builtin family id");
    if (kind) *kind = f->kind;
    if (nparams) *nparams = f->nparams;
    return FTS_OK;
}

extern "C" int64_t fts_evals_for_level(int D, int level)
{
    // cumulative evaluations (2n+1)^D with n = 2^(level+1)
    int64_t m = (int64_t(2) << (level + 1)) + 1;
    return D == 1 ? m : (D == 2 ? m * m : m * m * m);
}
extern "C" int64_t fts_evals_fixed(int D, int n)
{
    int64_t m = 2 * int64_t(n) + 1;
    return D == 1 ? m : (D == 2 ? m * m : m * m * m);
}

// ------------------------------------------------------------------------------- tablegen ----
static int check_dtype(int dtype)
{
    if (dtype != FTS_F32 && dtype != FTS_F64 && dtype != FTS_DD64) return set_error(FTS_ERR_INVALID_ARGUMENT, "unknown dtype");
    return FTS_OK;
}

extern "C" int fts_window(int dtype, int which, int D, void* tm_out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (D < 1 || which < 0 || which > 2 || !tm_out) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_window: bad arguments");
    return fts_tg_window(dtype, which, D, tm_out) ? set_error(FTS_ERR_INTERNAL, "fts_window failed") : FTS_OK;
}
extern "C" int fts_tanhsinh(int dtype, int N, int D, void* x_out, void* w_out, void* h_out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (N < 2 || D < 1) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_tanhsinh: N must be >= 2 and D >= 1");
    return fts_tg_fixed(dtype, N, D, x_out, w_out, h_out) ? set_error(FTS_ERR_INTERNAL, "fts_tanhsinh failed") : FTS_OK;
}
extern "C" int fts_hopt(int dtype, int N, double d, void* h_out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (N < 1 || !(d > 0) || !h_out) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_hopt: N must be >= 1 and d > 0");
    return fts_tg_hopt(dtype, N, d, h_out) ? set_error(FTS_ERR_INTERNAL, "fts_hopt failed") : FTS_OK;
}
extern "C" int fts_hmax(int dtype, int N, int D, void* h_out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (N < 2 || D < 1 || !h_out) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_hmax: N must be >= 2 and D >= 1");
    return fts_tg_hmax(dtype, N, D, h_out) ? set_error(FTS_ERR_INTERNAL, "fts_hmax failed") : FTS_OK;
}
extern "C" int fts_tanhsinh_opt(int dtype, int N, double d, int D, int64_t capacity, void* x_out, void* w_out, void* h_out, int64_t* count)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (N < 1 || D < 1 || !(d > 0) || capacity < 0 || !h_out || !count || (capacity > 0 && (!x_out || !w_out)))
        return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_tanhsinh_opt: bad arguments");
    const long k = fts_tg_opt(dtype, N, d, D, (long)capacity, x_out, w_out, h_out);
    if (k < 0) return set_error(FTS_ERR_INTERNAL, "fts_tanhsinh_opt failed");
    *count = k;
    return FTS_OK;
}
extern "C" int fts_cache1d_generate(int dtype, int max_levels, int complement, void* tm_out, void* initial_x, void* initial_w,
                                    void* initial_c, void* xs_flat, void* ws_flat, void* cs_flat)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (max_levels < 0) return set_error(FTS_ERR_INVALID_ARGUMENT, "`max_levels` must be nonnegative.");  // caches.jl:60
    if (max_levels > 28) return set_error(FTS_ERR_INVALID_ARGUMENT, "max_levels > 28 is not supported");
    return fts_tg_cache1d(dtype, max_levels, complement, tm_out, initial_x, initial_w, initial_c, xs_flat, ws_flat, cs_flat)
               ? set_error(FTS_ERR_INTERNAL, "cache generation failed") : FTS_OK;
}
extern "C" int fts_cache_tensor_generate(int dtype, int D, int max_levels, void* tm_out, void* initial_x, void* initial_w,
                                         void* xs_flat, void* ws_flat)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (max_levels < 0) return set_error(FTS_ERR_INVALID_ARGUMENT, "`max_levels` must be nonnegative.");  // caches.jl:149
    if (D != 2 && D != 3) return set_error(FTS_ERR_INVALID_ARGUMENT, "tensor caches exist for D = 2, 3");
    if (max_levels > 24) return set_error(FTS_ERR_INVALID_ARGUMENT, "max_levels > 24 is not supported");
    return fts_tg_cache_tensor(dtype, D, max_levels, tm_out, initial_x, initial_w, xs_flat, ws_flat)
               ? set_error(FTS_ERR_INTERNAL, "cache generation failed") : FTS_OK;
}
extern "C" int fts_cache_tensor_cmpl_generate(int dtype, int max_levels, void* tm_out, void* initial_x, void* initial_w,
                                              void* initial_c, void* xs_flat, void* ws_flat, void* cs_flat)
{
    if (dtype != FTS_F32 && dtype != FTS_F64) return set_error(FTS_ERR_UNSUPPORTED, "2D complement tables: f32/f64 only");
    if (max_levels < 0 || max_levels > 24) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad max_levels");
    // window t_w_max(T, 2): products of two weights/complements stay above floatmin (core.jl:207-221)
    unsigned char tm[16];
    fts_tg_window(dtype, 2, 2, tm);
    memcpy(tm_out, tm, dtype_size(dtype));
    if (fts_tg_tensor_complements(dtype, max_levels, tm, initial_x, initial_w, initial_c, xs_flat, ws_flat, cs_flat))
        return set_error(FTS_ERR_INTERNAL, "tensor complement table generation failed");
    return FTS_OK;
}

// ------------------------------------------------------------------------------- uploads -----
template <class T>
static int upload_nodes(const T* x, const T* w, const T* c, long long count, void** d_nodes, void** d_nodes_c)
{
    *d_nodes = nullptr;
    *d_nodes_c = nullptr;
    if (count == 0) return FTS_OK;
    std::vector<fts::Node<T>> h(count);
    for (long long i = 0; i < count; ++i) { h[i].x = x[i]; h[i].w = w[i]; }
    FTS_CUDA(cudaMalloc(d_nodes, sizeof(fts::Node<T>) * count));
    FTS_CUDA(cudaMemcpy(*d_nodes, h.data(), sizeof(fts::Node<T>) * count, cudaMemcpyHostToDevice));
    if (c) {
        std::vector<fts::NodeC<T>> hc(count);
        for (long long i = 0; i < count; ++i) { hc[i].x = x[i]; hc[i].w = w[i]; hc[i].c = c[i]; hc[i].pad = T(0.0); }
        FTS_CUDA(cudaMalloc(d_nodes_c, sizeof(fts::NodeC<T>) * count));
        FTS_CUDA(cudaMemcpy(*d_nodes_c, hc.data(), sizeof(fts::NodeC<T>) * count, cudaMemcpyHostToDevice));
    }
    return FTS_OK;
}

static int upload_nodes_any(int dtype, const void* x, const void* w, const void* c, long long count, void** dn, void** dc)
{
    switch (dtype) {
    case FTS_F32: return upload_nodes<float>((const float*)x, (const float*)w, (const float*)c, count, dn, dc);
    case FTS_F64: return upload_nodes<double>((const double*)x, (const double*)w, (const double*)c, count, dn, dc);
    case FTS_DD64: return upload_nodes<fts::dd>((const fts::dd*)x, (const fts::dd*)w, (const fts::dd*)c, count, dn, dc);
    }
    return set_error(FTS_ERR_INVALID_ARGUMENT, "unknown dtype");
}

extern "C" int fts_tables_fixed_upload(int dtype, int n, const void* x, const void* w, const void* h, fts_tables* out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (n < 0 || !x || !w || !h || !out) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_tables_fixed_upload: bad arguments");
    if (int rc = ensure_init()) return rc;
    fts_tables_s* t = new fts_tables_s();
    t->dtype = dtype;
    t->n = n;
    memset(t->h, 0, sizeof t->h);
    memcpy(t->h, h, dtype_size(dtype));
    void* dummy = nullptr;
    int rc = upload_nodes_any(dtype, x, w, nullptr, n, &t->d_grid, &dummy);
    if (rc) { delete t; return rc; }
    *out = t;
    return FTS_OK;
}
extern "C" int fts_tables_free(fts_tables t)
{
    if (!t) return FTS_OK;
    if (t->d_grid) cudaFree(t->d_grid);
    delete t;
    return FTS_OK;
}

static int make_cache(int dtype, int D, int kind, int max_levels, const void* tm, const void* ix, const void* iw, const void* ic,
                      const void* xs, const void* ws, const void* cs, fts_cache* out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (max_levels < 0) return set_error(FTS_ERR_INVALID_ARGUMENT, "`max_levels` must be nonnegative.");
    if (!tm || !ix || !iw || !out || (max_levels > 0 && (!xs || !ws))) return set_error(FTS_ERR_INVALID_ARGUMENT, "cache upload: null argument");
    if (kind == 1 && (!ic || (max_levels > 0 && !cs))) return set_error(FTS_ERR_INVALID_ARGUMENT, "complement cache upload needs initial_c and cs");
    if (int rc = ensure_init()) return rc;
    const size_t es = dtype_size(dtype);
    fts_cache_s* c = new fts_cache_s();
    c->dtype = dtype; c->D = D; c->kind = kind; c->max_levels = max_levels;
    c->count = D == 0 ? ((2LL << max_levels) - 2) : ((4LL << max_levels) - 4);
    memset(c->tm, 0, sizeof c->tm); memset(c->ix, 0, sizeof c->ix); memset(c->iw, 0, sizeof c->iw); memset(c->ic, 0, sizeof c->ic);
    memcpy(c->tm, tm, es); memcpy(c->ix, ix, 2 * es); memcpy(c->iw, iw, 2 * es);
    if (ic) memcpy(c->ic, ic, 2 * es);
    int rc = upload_nodes_any(dtype, xs, ws, kind == 1 ? cs : nullptr, c->count, &c->d_nodes, &c->d_nodes_c);
    if (rc) { delete c; return rc; }
    *out = c;
    return FTS_OK;
}

extern "C" int fts_cache1d_upload(int dtype, int kind, int max_levels, const void* tm, const void* initial_x, const void* initial_w,
                                  const void* initial_c, const void* xs_flat, const void* ws_flat, const void* cs_flat, fts_cache* out)
{
    if (kind != 0 && kind != 1) return set_error(FTS_ERR_INVALID_ARGUMENT, "cache kind must be 0 (regular) or 1 (complement)");
    return make_cache(dtype, 0, kind, max_levels, tm, initial_x, initial_w, initial_c, xs_flat, ws_flat, cs_flat, out);
}
extern "C" int fts_cache_tensor_upload(int dtype, int D, int max_levels, const void* tm, const void* initial_x, const void* initial_w,
                                       const void* xs_flat, const void* ws_flat, fts_cache* out)
{
    if (D != 2 && D != 3) return set_error(FTS_ERR_INVALID_ARGUMENT, "tensor caches exist for D = 2, 3");
    return make_cache(dtype, D, 0, max_levels, tm, initial_x, initial_w, nullptr, xs_flat, ws_flat, nullptr, out);
}
extern "C" int fts_cache_tensor_cmpl_upload(int dtype, int max_levels, const void* tm, const void* initial_x, const void* initial_w,
                                            const void* initial_c, const void* xs_flat, const void* ws_flat, const void* cs_flat, fts_cache* out)
{
    return make_cache(dtype, 2, 1, max_levels, tm, initial_x, initial_w, initial_c, xs_flat, ws_flat, cs_flat, out);
}
// ---- on-device generation and download ----------------------------------------
template <class T> static int tables_create_dev(int N, int D, fts_tables_s* t)
{
    const int n = N / 2;  // core.jl:257
    T tm;
    fts_tg_window(t->dtype, 0, D, &tm);
    const T h = tm / T((double)n);
    memset(t->h, 0, sizeof t->h);
    memcpy(t->h, &h, sizeof(T));
    t->n = n;
    t->d_grid = nullptr;
    if (n == 0) return FTS_OK;
    FTS_CUDA(cudaMalloc(&t->d_grid, sizeof(fts::Node<T>) * n));
    fts::k_tablegen_fixed<T><<<(n + 127) / 128, 128, 0, g_stream>>>((fts::Node<T>*)t->d_grid, n, h, tm);
    FTS_CUDA(cudaGetLastError());
    FTS_CUDA(cudaStreamSynchronize(g_stream));
    return FTS_OK;
}

extern "C" int fts_tables_fixed_create(int dtype, int N, int D, fts_tables* out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (N < 2 || D < 1 || !out) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_tables_fixed_create: N must be >= 2 and D >= 1");
    if (int rc = ensure_init()) return rc;
    fts_tables_s* t = new fts_tables_s();
    t->dtype = dtype;
    int rc = dtype == FTS_F32 ? tables_create_dev<float>(N, D, t) : dtype == FTS_F64 ? tables_create_dev<double>(N, D, t) : tables_create_dev<fts::dd>(N, D, t);
    if (rc) { if (t->d_grid) cudaFree(t->d_grid); delete t; return rc; }
    *out = t;
    return FTS_OK;
}

template <class T> static int cache_create_dev(fts_cache_s* c, const void* tm_in)
{
    T tm;
    memcpy(&tm, tm_in, sizeof(T));
    memset(c->tm, 0, sizeof c->tm); memset(c->ix, 0, sizeof c->ix); memset(c->iw, 0, sizeof c->iw); memset(c->ic, 0, sizeof c->ic);
    memcpy(c->tm, &tm, sizeof(T));
    T* d_init = nullptr;
    FTS_CUDA(cudaMalloc((void**)&d_init, 6 * sizeof(T)));
    fts::k_tablegen_initial<T><<<1, 32, 0, g_stream>>>(d_init, tm);
    T init[6];
    cudaError_t e = cudaMemcpyAsync(init, d_init, sizeof init, cudaMemcpyDeviceToHost, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    cudaFree(d_init);
    FTS_CUDA(e);
    memcpy(c->ix, init, 2 * sizeof(T)); memcpy(c->iw, init + 2, 2 * sizeof(T));
    if (c->kind == 1) memcpy(c->ic, init + 4, 2 * sizeof(T));
    c->d_nodes = nullptr; c->d_nodes_c = nullptr;
    if (c->count == 0) return FTS_OK;
    FTS_CUDA(cudaMalloc(&c->d_nodes, sizeof(fts::Node<T>) * c->count));
    if (c->kind == 1) FTS_CUDA(cudaMalloc(&c->d_nodes_c, sizeof(fts::NodeC<T>) * c->count));
    const unsigned blocks = (unsigned)((c->count + 127) / 128);
    fts::k_tablegen_levels<T><<<blocks, 128, 0, g_stream>>>((fts::Node<T>*)c->d_nodes, (fts::NodeC<T>*)c->d_nodes_c, c->count, tm, c->D != 0);
    FTS_CUDA(cudaGetLastError());
    FTS_CUDA(cudaStreamSynchronize(g_stream));
    return FTS_OK;
}

// D = 0: 1D cache (kind 0 regular / 1 complement); D = 2, 3: tensor cache; D = 2 with kind 1: the
// 2D complement table (window t_w_max(T, 2), see fts_cache_tensor_cmpl_generate)
extern "C" int fts_cache_create(int dtype, int D, int kind, int max_levels, fts_cache* out)
{
    if (int rc = check_dtype(dtype)) return rc;
    if (!out) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_cache_create: null output");
    if (max_levels < 0) return set_error(FTS_ERR_INVALID_ARGUMENT, "`max_levels` must be nonnegative.");  // caches.jl:60,149
    if (max_levels > 28) return set_error(FTS_ERR_INVALID_ARGUMENT, "max_levels > 28 is not supported");
    if ((D != 0 && D != 2 && D != 3) || (kind != 0 && kind != 1) || (kind == 1 && D == 3))
        return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_cache_create: D must be 0 (1D), 2 or 3; complement tables exist for D = 0 and 2");
    if (int rc = ensure_init()) return rc;
    unsigned char tm[16];
    memset(tm, 0, sizeof tm);
    // windows: caches.jl:47-48 (1D), :152 (tensor); scalar Newton iterations stay on the host
    if (D == 0) fts_tg_window(dtype, kind == 1 ? 2 : 0, 1, tm);
    else if (kind == 1) fts_tg_window(dtype, 2, 2, tm);
    else fts_tg_window(dtype, 0, D, tm);
    fts_cache_s* c = new fts_cache_s();
    c->dtype = dtype; c->D = D; c->kind = kind; c->max_levels = max_levels;
    c->count = D == 0 ? ((2LL << max_levels) - 2) : ((4LL << max_levels) - 4);
    int rc = dtype == FTS_F32 ? cache_create_dev<float>(c, tm) : dtype == FTS_F64 ? cache_create_dev<double>(c, tm) : cache_create_dev<fts::dd>(c, tm);
    if (rc) { if (c->d_nodes) cudaFree(c->d_nodes); if (c->d_nodes_c) cudaFree(c->d_nodes_c); delete c; return rc; }
    *out = c;
    return FTS_OK;
}

template <class T> static int download_nodes(const void* d_nodes, const void* d_nodes_c, long long count, void* x, void* w, void* cc)
{
    if (count == 0) return FTS_OK;
    if (d_nodes_c && cc) {
        std::vector<fts::NodeC<T>> h(count);
        FTS_CUDA(cudaMemcpy(h.data(), d_nodes_c, sizeof(fts::NodeC<T>) * count, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < count; ++i) { if (x) ((T*)x)[i] = h[i].x; if (w) ((T*)w)[i] = h[i].w; ((T*)cc)[i] = h[i].c; }
        return FTS_OK;
    }
    std::vector<fts::Node<T>> h(count);
    FTS_CUDA(cudaMemcpy(h.data(), d_nodes, sizeof(fts::Node<T>) * count, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < count; ++i) { if (x) ((T*)x)[i] = h[i].x; if (w) ((T*)w)[i] = h[i].w; }
    return FTS_OK;
}

// the flat level arrays of a cache, in the layout fts_cache*_upload takes (any pointer may be null)
extern "C" int fts_cache_download(fts_cache c, void* tm, void* initial_x, void* initial_w, void* initial_c, void* xs_flat, void* ws_flat, void* cs_flat)
{
    if (!c) return set_error(FTS_ERR_INVALID_ARGUMENT, "null cache");
    if (cs_flat && c->kind != 1) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_cache_download: not a complement cache");
    const size_t es = dtype_size(c->dtype);
    if (tm) memcpy(tm, c->tm, es);
    if (initial_x) memcpy(initial_x, c->ix, 2 * es);
    if (initial_w) memcpy(initial_w, c->iw, 2 * es);
    if (initial_c) memcpy(initial_c, c->ic, 2 * es);
    switch (c->dtype) {
    case FTS_F32: return download_nodes<float>(c->d_nodes, c->d_nodes_c, c->count, xs_flat, ws_flat, cs_flat);
    case FTS_F64: return download_nodes<double>(c->d_nodes, c->d_nodes_c, c->count, xs_flat, ws_flat, cs_flat);
    case FTS_DD64: return download_nodes<fts::dd>(c->d_nodes, c->d_nodes_c, c->count, xs_flat, ws_flat, cs_flat);
    }
    return set_error(FTS_ERR_INTERNAL, "bad dtype");
}
extern "C" int fts_tables_download(fts_tables t, int* n, void* x, void* w, void* h)
{
    if (!t) return set_error(FTS_ERR_INVALID_ARGUMENT, "null tables");
    if (n) *n = t->n;
    if (h) memcpy(h, t->h, dtype_size(t->dtype));
    if (!x && !w) return FTS_OK;
    switch (t->dtype) {
    case FTS_F32: return download_nodes<float>(t->d_grid, nullptr, t->n, x, w, nullptr);
    case FTS_F64: return download_nodes<double>(t->d_grid, nullptr, t->n, x, w, nullptr);
    case FTS_DD64: return download_nodes<fts::dd>(t->d_grid, nullptr, t->n, x, w, nullptr);
    }
    return set_error(FTS_ERR_INTERNAL, "bad dtype");
}

extern "C" int fts_cache_info(fts_cache c, int* dtype, int* D, int* kind, int* max_levels)
{
    if (!c) return set_error(FTS_ERR_INVALID_ARGUMENT, "null cache");
    if (dtype) *dtype = c->dtype;
    if (D) *D = c->D;
    if (kind) *kind = c->kind;
    if (max_levels) *max_levels = c->max_levels;
    return FTS_OK;
}
extern "C" int fts_cache_free(fts_cache c)
{
    if (!c) return FTS_OK;
    if (c->d_nodes) cudaFree(c->d_nodes);
    if (c->d_nodes_c) cudaFree(c->d_nodes_c);
    delete c;
    return FTS_OK;
}

// ------------------------------------------------------------------------------- integrands --
extern "C" int fts_integrand_builtin(int family, fts_integrand* out)
{
    const FamilyInfo* fi = family_info(family);
    if (!fi || !out) return set_error(FTS_ERR_INVALID_ARGUMENT, "unknown builtin family id");
    fts_integrand_s* f = new fts_integrand_s();
    f->builtin = 1; f->family = family; f->kind = fi->kind; f->nparams = fi->nparams;
    f->module = nullptr;
    *out = f;
    return FTS_OK;
}
extern "C" int fts_integrand_info(fts_integrand f, int* kind, int* nparams, int* is_builtin)
{
    if (!f) return set_error(FTS_ERR_INVALID_ARGUMENT, "null integrand");
    if (kind) *kind = f->kind;
    if (nparams) *nparams = f->nparams;
    if (is_builtin) *is_builtin = f->builtin;
    return FTS_OK;
}
extern "C" int fts_integrand_free(fts_integrand f)
{
    delete f;  // NVRTC modules stay in the process-wide compile cache
    return FTS_OK;
}

static const void* integrand_kernel(fts_integrand f, int dtype, int kid, int fast)
{
    if (f->builtin) return find_kernel(f->family, dtype, kid, fast);
    return fts_nvrtc_kernel(f->module, dtype, kid, fast);
}

// device-side alias of a pinned, mapped host allocation (cudaHostAlloc / cudaHostRegister under
// unified addressing); nullptr for pageable or foreign memory
static void* mapped_dev_ptr(const void* host)
{
    if (!host) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// Pinned host buffers are used IN PLACE: the kernels read the per-integral inputs and write the
// results straight through the caller's memory over PCIe, overlapped with the arithmetic CTA by
// CTA, and no staging copy is queued (config B, 8 MiB in + 8 MiB out: 0.33 ms against 0.43 ms for
// the staged 4-chunk pipeline and 0.29 ms for the kernel alone).
// FTS_B200_ZEROCOPY (bit mask, default 3): 1 = inputs in place, 2 = results in place; 0 = stage
// everything through the chunk pipeline.
static int zero_copy_mode()
{
    static const int mode = [] { const char* e = getenv("FTS_B200_ZEROCOPY"); return e ? (atoi(e) & 3) : 3; }();
    return mode;
}
static bool input_in_place() { return (zero_copy_mode() & 1) != 0; }
static bool output_in_place() { return (zero_copy_mode() & 2) != 0; }

// ------------------------------------------------------------------------------- launches ----
template <class T> static T load_elem(const void* p) { T v; memcpy(&v, p, sizeof(T)); return v; }
static double elem_hi(int dtype, const void* p) { return dtype == FTS_F32 ? (double)load_elem<float>(p) : load_elem<double>(p); }

// dynamic shared memory a kernel may ask for: the opt-in maximum minus what the kernels declare
// statically (reduction scratch, flags, the 16 KiB exp table region / 3 KiB log table)
constexpr size_t STATIC_SMEM_MAX = 20 * 1024;
static size_t dyn_smem_limit() { return (size_t)g_prop.sharedMemPerBlockOptin - STATIC_SMEM_MAX; }

// opt-in dynamic shared memory above 48 KiB: set once per kernel and size (the attribute call costs a
// microsecond or two, which a 30 us launch notices)
static int ensure_dyn_smem(const void* fn, size_t smem)
{
    if (smem + STATIC_SMEM_MAX <= 48 * 1024) return FTS_OK;  // static + dynamic within the default limit: no opt-in needed
    if (smem > dyn_smem_limit())
        return set_error(FTS_ERR_UNSUPPORTED, "level needs " + std::to_string(smem) + " B of shared memory per CTA (max " +
                                                  std::to_string(dyn_smem_limit()) + ")");
    thread_local std::vector<std::pair<const void*, size_t>> done;  // per host thread: no lock on the launch path
    for (auto& e : done)
        if (e.first == fn) {
            if (e.second >= smem) return FTS_OK;
            break;
        }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        FTS_CUDA(cudaKernelSetAttributeForDevice((cudaKernel_t)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem, g_device));
    }
    for (auto& d : done)
        if (d.first == fn) { d.second = smem; return FTS_OK; }
    done.emplace_back(fn, smem);
    return FTS_OK;
}

static int launch_raw(const void* fn, unsigned grid, int block, size_t smem, void* args, cudaStream_t s)
{
    if (int rc = ensure_dyn_smem(fn, smem)) return rc;
    void* params[] = {args};
    FTS_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(block), params, smem, s));
    g_launches.fetch_add(1);
    return FTS_OK;
}

// Second kernel of a two-kernel call (the queue passes k_adaptive1d_tail / k_adaptive2d_cta): launched with
// programmatic stream serialisation, so its CTAs are set up while the first kernel's last wave still runs
// and wait in `griddepcontrol.wait` for the first kernel's completion and memory flush — the launch
// latency of the second kernel (4-8 us on an idle stream) disappears behind the first.
static int launch_dependent_params(const void* fn, unsigned grid, int block, void** params, cudaStream_t s)
{
    static const bool use_pdl = [] { const char* e = getenv("FTS_B200_PDL"); return !e || atoi(e) != 0; }();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    FTS_CUDA(cudaLaunchKernelExC(&cfg, fn, params));
    g_launches.fetch_add(1);
    return FTS_OK;
}

static int launch_dependent(const void* fn, unsigned grid, int block, size_t smem, void* args, cudaStream_t s)
{
    static const bool use_pdl = [] { const char* e = getenv("FTS_B200_PDL"); return !e || atoi(e) != 0; }();
    if (!use_pdl) return launch_raw(fn, grid, block, smem, args, s);
    if (int rc = ensure_dyn_smem(fn, smem)) return rc;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    void* params[] = {args};
    FTS_CUDA(cudaLaunchKernelExC(&cfg, fn, params));
    g_launches.fetch_add(1);
    return FTS_OK;
}

// one cluster of `csize` CTAs (thread-block cluster, distributed shared memory) per work item
static int launch_cluster(const void* fn, unsigned nclusters, unsigned csize, int block, size_t smem, void* args, cudaStream_t s)
{
    if (int rc = ensure_dyn_smem(fn, smem)) return rc;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(nclusters * csize);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    void* params[] = {args};
    FTS_CUDA(cudaLaunchKernelExC(&cfg, fn, params));
    g_launches.fetch_add(1);
    return FTS_OK;
}

// every CTA resident at once (the kernel has a grid-wide barrier): `grid` is cut to what the device can hold
static int launch_cooperative(const void* fn, unsigned grid, int block, size_t smem, void* args, cudaStream_t s)
{
    if (int rc = ensure_dyn_smem(fn, smem)) return rc;
    void* params[] = {args};
    for (int per_sm = 4; per_sm >= 1; --per_sm) {
        const unsigned cap = (unsigned)g_prop.multiProcessorCount * per_sm;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid < cap ? grid : cap);
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeCooperative;
        attr.val.cooperative = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelExC(&cfg, fn, params);
        if (e == cudaSuccess) {
            g_launches.fetch_add(1);
            return FTS_OK;
        }
        cudaGetLastError();
        if (e != cudaErrorCooperativeLaunchTooLarge || per_sm == 1) return set_error(FTS_ERR_CUDA, std::string("cooperative launch: ") + cudaGetErrorString(e));
    }
    return FTS_OK;
}

// threads per CTA of the 1D thread-per-integral kernels: 128 (8 192 CTAs for 2^20 Float64
// integrals: shorter tail than 256, 0.286 vs 0.290 ms on config B); Double64 prefers 256 (config E
// 0.378 vs 0.394 ms); smaller CTAs for small batches were measured and do not help.
// FTS_B200_TPI_BLOCK pins the value (experiments; must not exceed the kernels' __launch_bounds__).
static int tpi_block(int dim, int dtype)
{
    static const int pinned = [] {
        const char* e = getenv("FTS_B200_TPI_BLOCK");
        const int v = e ? atoi(e) : 0;
        return (v >= 32 && v <= 256 && v % 32 == 0) ? v : 0;
    }();
    if (dim != 1) return 128;
    if (pinned) return pinned;
    return dtype == FTS_DD64 ? 256 : 128;
}

static int launch(const void* fn, long long threads, int block, void* args, cudaStream_t s)
{
    if (threads <= 0) return FTS_OK;
    long long blocks = (threads + block - 1) / block;
    if (blocks > 0x7fffffffLL) return set_error(FTS_ERR_INVALID_ARGUMENT, "batch too large for one launch");
    return launch_raw(fn, (unsigned)blocks, block, 0, args, s);
}

// Which kernel style serves a launch.  Thread-per-integral needs enough integrals
// to fill 148 SMs; below that one CTA per integral; below ~2 CTAs per SM and with deep levels
// (2D/3D) one whole-grid launch per level.
enum Style { STYLE_TPI, STYLE_CTA, STYLE_GRID };
static Style pick_style(int order, int dim, long long batch, int max_levels, bool have_cta, bool have_grid)
{
    if (order == FTS_ORDER_REFERENCE) return STYLE_TPI;
    const long long sms = g_prop.multiProcessorCount;
    const long long tpi_min = dim == 1 ? 32 * sms : 64 * sms;  // >= 4.7k / 9.5k integrals
    if (order == FTS_ORDER_AUTO && batch >= tpi_min) return STYLE_TPI;
    if (order == FTS_ORDER_FAST && batch >= 8 * tpi_min) return STYLE_TPI;
    if (dim >= 2 && have_grid && batch < 2 * sms && max_levels >= (dim == 3 ? 4 : 7)) return STYLE_GRID;
    return have_cta ? STYLE_CTA : STYLE_TPI;
}

template <class T> static void fill_cache_args(fts::Args<T>& a, fts_cache c)
{
    a.nodes = (const fts::Node<T>*)c->d_nodes;
    a.nodes_c = (const fts::NodeC<T>*)c->d_nodes_c;
    a.tm = load_elem<T>(c->tm);
    for (int i = 0; i < 2; ++i) {
        a.ix[i] = load_elem<T>(c->ix + i * sizeof(T));
        a.iw[i] = load_elem<T>(c->iw + i * sizeof(T));
        a.ic[i] = load_elem<T>(c->ic + i * sizeof(T));
    }
}

// generic n-D tensor product (kinds FTS_KIND_ND + D): gridDim.x CTAs per integral, sized to the (2n+1)^D points
template <class T>
static int run_fixed_nd(fts_tables t, fts_integrand f, const void* low, const void* up, int bstride, const void* params, int pstride,
                        long long batch, void* out_value, cudaStream_t s)
{
    if (batch == 0) return FTS_OK;
    if (batch > 65535) return set_error(FTS_ERR_INVALID_ARGUMENT, "n-D tensor products: at most 65535 integrals per call");
    const int dim = kind_dim(f->kind);
    const void* fn = integrand_kernel(f, t->dtype, KID_FIXED_ND, 1);
    if (!fn) return set_error(FTS_ERR_UNSUPPORTED, "n-D tensor products: float32 / float64 NVRTC integrands");
    const double m = 2.0 * t->n + 1.0;
    double points = 1.0;
    for (int d = 0; d < dim; ++d) points *= m;
    if (points > 9.0e18) return set_error(FTS_ERR_INVALID_ARGUMENT, "n-D tensor product: (2N+1)^D exceeds 2^63 points");
    // >= 2 runs of up to 32 points per thread, at most 4 CTAs per SM over the whole batch
    double want = points / m * ceil(m / 32.0) / (fts::CTA_THREADS * 2.0);
    const double cap = (double)g_prop.multiProcessorCount * 4.0 / (double)batch;
    if (want > cap) want = cap;
    if (want < 1.0) want = 1.0;
    const unsigned blocks = (unsigned)want;
    T* scratch = nullptr;
    const size_t part = 2 * (size_t)blocks * batch;
    FTS_CUDA(cudaMallocAsync((void**)&scratch, part * sizeof(T) + (size_t)batch * sizeof(unsigned int), s));
    unsigned int* ticket = (unsigned int*)(scratch + part);
    FTS_CUDA(cudaMemsetAsync(ticket, 0, (size_t)batch * sizeof(unsigned int), s));
    fts::Args<T> a;
    memset(&a, 0, sizeof a);
    a.grid = (const fts::Node<T>*)t->d_grid;
    a.n = t->n;
    a.h = load_elem<T>(t->h);
    a.low = (const T*)low; a.up = (const T*)up; a.bstride = bstride;
    a.params = (const T*)params; a.pstride = pstride;
    a.batch = batch;
    a.value = (T*)out_value;
    a.block_partials = scratch;
    a.ticket = ticket;
    const size_t smem = 2 * (size_t)(2 * t->n + 1) * sizeof(T);
    if (int rc = ensure_dyn_smem(fn, smem)) return rc;
    void* kargs[] = {&a};
    FTS_CUDA(cudaLaunchKernel(fn, dim3(blocks, (unsigned)batch), dim3(fts::CTA_THREADS), kargs, smem, s));
    g_launches.fetch_add(1);
    FTS_CUDA(cudaFreeAsync(scratch, s));
    return FTS_OK;
}

template <class T>
static int run_fixed(fts_tables t, fts_integrand f, int order, const void* low, const void* up, int bstride, const void* params,
                     int pstride, long long batch, void* out_value, cudaStream_t s)
{
    const int dim = kind_dim(f->kind);
    if (f->kind > FTS_KIND_ND) {
        if constexpr (sizeof(T) == 16) return set_error(FTS_ERR_UNSUPPORTED, "n-D tensor products: float32 / float64");
        else return run_fixed_nd<T>(t, f, low, up, bstride, params, pstride, batch, out_value, s);
    }
    const void* fn_cta = t->dtype == FTS_DD64 ? nullptr : integrand_kernel(f, t->dtype, KID_FIXED_CTA, 1);
    const size_t smem = dim == 1 ? 0 : (size_t)(dim == 2 ? 5 : fts::SM3_ARRAYS) * t->n * sizeof(T);
    const bool cta_ok = fn_cta && smem <= dyn_smem_limit();
    Style st = pick_style(order, dim, batch, 0, cta_ok, false);
    fts::Args<T> a;
    memset(&a, 0, sizeof a);
    a.grid = (const fts::Node<T>*)t->d_grid;
    a.n = t->n;
    a.h = load_elem<T>(t->h);
    a.low = (const T*)low; a.up = (const T*)up; a.bstride = bstride;
    a.params = (const T*)params; a.pstride = pstride;
    a.batch = batch;
    a.value = (T*)out_value;
    if (batch == 0) return FTS_OK;
    if (st == STYLE_CTA) return launch_raw(fn_cta, (unsigned)batch, fts::CTA_THREADS, smem, &a, s);
    // REFERENCE order on a few 2D/3D integrals: rows in parallel, folded in order (k_fixed2d_ord)
    // (1D: only long grids — a thread adds 2 x 256 evaluations faster than a CTA can be scheduled)
    if (order == FTS_ORDER_REFERENCE && (dim >= 2 || t->n >= 256) && t->dtype != FTS_DD64 && batch < 64LL * g_prop.multiProcessorCount) {
        static const bool use_ord = [] { const char* e = getenv("FTS_B200_ORDERED"); return !e || atoi(e) != 0; }();
        const void* fn_ord = use_ord ? integrand_kernel(f, t->dtype, KID_FIXED_ORD, 0) : nullptr;
        if (fn_ord) {
            unsigned csize = 1;  // a cluster of CTAs per integral when the integrals are too few to occupy the SMs
            while (csize < 8 && batch * (long long)csize * 2 <= 2LL * g_prop.multiProcessorCount) csize *= 2;
            a.level = 4096;  // terms per tile buffer: with 8 x 256 threads every thread gets a 3D row per tile
            return launch_cluster(fn_ord, (unsigned)batch, csize, 256, 2 * (size_t)a.level * sizeof(T), &a, s);
        }
        if (use_ord) g_last_error.clear();
    }
    const int fast = order == FTS_ORDER_REFERENCE ? 0 : (order == FTS_ORDER_FAST ? 1 : 0);
    const void* fn = integrand_kernel(f, t->dtype, KID_FIXED_TPI, fast);
    if (!fn) return set_error(FTS_ERR_UNSUPPORTED, "no fixed-grid kernel built for this integrand/dtype/order");
    return launch(fn, batch, tpi_block(dim, t->dtype), &a, s);
}

static int check_fixed(fts_tables t, fts_integrand f, int order, const void* low, const void* up, int bstride, const void* params,
                       int pstride, long long batch, void* out_value)
{
    if (!t || !f) return set_error(FTS_ERR_INVALID_ARGUMENT, "null handle");
    if (kind_cmpl(f->kind)) return set_error(FTS_ERR_UNSUPPORTED, "fixed-grid complement integration is not on the reference's exported path");
    const int dim = kind_dim(f->kind);
    if (bstride != 0 && bstride != dim) return set_error(FTS_ERR_DIMENSION_MISMATCH, "low/up must have length " + std::to_string(dim) + " per integral");
    if (f->nparams > 0 && (!params || (pstride != 0 && pstride < f->nparams))) return set_error(FTS_ERR_INVALID_ARGUMENT, "integrand needs " + std::to_string(f->nparams) + " parameters per integral");
    if (batch < 0 || !low || !up || (batch > 0 && !out_value)) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad batch arguments");
    if (order != FTS_ORDER_REFERENCE && order != FTS_ORDER_FAST && order != FTS_ORDER_AUTO) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad order");
    return ensure_init();
}

extern "C" int fts_integrate_batch_dev(fts_tables t, fts_integrand f, int order, const void* low, const void* up, int bounds_stride,
                                       const void* params, int params_stride, int64_t batch, void* out_value, void* stream)
{
    if (int rc = check_fixed(t, f, order, low, up, bounds_stride, params, params_stride, batch, out_value)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    switch (t->dtype) {
    case FTS_F32: return run_fixed<float>(t, f, order, low, up, bounds_stride, params, params_stride, batch, out_value, s);
    case FTS_F64: return run_fixed<double>(t, f, order, low, up, bounds_stride, params, params_stride, batch, out_value, s);
    case FTS_DD64: return run_fixed<fts::dd>(t, f, order, low, up, bounds_stride, params, params_stride, batch, out_value, s);
    }
    return set_error(FTS_ERR_INTERNAL, "bad dtype");
}

// f(x_b) for a batch of points: pointwise values of a 1D integrand (host buffers; synchronous)
template <class T> static int run_eval(const void* fn, fts_integrand f, const void* x, const void* params, int pstride, long long n, void* out)
{
    HostCtx* hc = nullptr;
    if (int rc = host_ctx(&hc)) return rc;
    cudaStream_t g_stream = hc->stream;  // this thread's stream
    const size_t np = f->nparams ? (pstride ? (size_t)n * pstride : (size_t)f->nparams) : 0;
    T *dx = nullptr, *dp = nullptr, *dv = nullptr;
    FTS_CUDA(cudaMalloc((void**)&dx, sizeof(T) * n));
    cudaError_t e = cudaMalloc((void**)&dv, sizeof(T) * n);
    if (e == cudaSuccess && np) e = cudaMalloc((void**)&dp, sizeof(T) * np);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, sizeof(T) * n, cudaMemcpyHostToDevice, g_stream);
    if (e == cudaSuccess && np) e = cudaMemcpyAsync(dp, params, sizeof(T) * np, cudaMemcpyHostToDevice, g_stream);
    int rc = FTS_OK;
    if (e == cudaSuccess) {
        fts::Args<T> a;
        memset(&a, 0, sizeof a);
        a.low = dx; a.up = dx; a.bstride = 1; a.params = dp; a.pstride = pstride; a.batch = n; a.value = dv;
        rc = launch(fn, n, 256, &a, g_stream);
        if (rc == FTS_OK) e = cudaMemcpyAsync(out, dv, sizeof(T) * n, cudaMemcpyDeviceToHost, g_stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    cudaFree(dx); cudaFree(dv); cudaFree(dp);
    if (rc) return rc;
    FTS_CUDA(e);
    return FTS_OK;
}

extern "C" int fts_eval_batch(fts_integrand f, int dtype, const void* x, const void* params, int params_stride, int64_t n, void* out_value)
{
    if (!f || !x || (n > 0 && !out_value) || n < 0) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_eval_batch: bad arguments");
    if (f->kind != FTS_KIND_1D) return set_error(FTS_ERR_UNSUPPORTED, "fts_eval_batch: integrands of kind f(x) only");
    if (dtype != FTS_F32 && dtype != FTS_F64) return set_error(FTS_ERR_UNSUPPORTED, "fts_eval_batch: f32/f64");
    if (f->nparams > 0 && (!params || (params_stride != 0 && params_stride < f->nparams))) return set_error(FTS_ERR_INVALID_ARGUMENT, "integrand needs " + std::to_string(f->nparams) + " parameters per point");
    if (int rc = ensure_init()) return rc;
    if (n == 0) return FTS_OK;
    const void* fn = integrand_kernel(f, dtype, KID_EVAL, 1);
    if (!fn) return set_error(FTS_ERR_UNSUPPORTED, "no pointwise-evaluation kernel for this integrand/dtype");
    return dtype == FTS_F32 ? run_eval<float>(fn, f, x, params, params_stride, n, out_value) : run_eval<double>(fn, f, x, params, params_stride, n, out_value);
}

// host <-> device staging for the host-buffer entry points
static int stage_in(HostCtx* hc, int slot, const void* host, size_t bytes, void** dev)
{
    *dev = nullptr;
    if (!host || bytes == 0) return FTS_OK;
    FTS_CUDA(hc->scratch[slot].need(bytes));
    FTS_CUDA(cudaMemcpyAsync(hc->scratch[slot].p, host, bytes, cudaMemcpyHostToDevice, hc->stream));
    *dev = hc->scratch[slot].p;
    return FTS_OK;
}
static int stage_out(HostCtx* hc, int slot, bool wanted, size_t bytes, void** dev)
{
    *dev = nullptr;
    if (!wanted || bytes == 0) return FTS_OK;
    FTS_CUDA(hc->scratch[slot].need(bytes));
    *dev = hc->scratch[slot].p;
    return FTS_OK;
}

extern "C" int fts_integrate_batch(fts_tables t, fts_integrand f, int order, const void* low, const void* up, int bounds_stride,
                                   const void* params, int params_stride, int64_t batch, void* out_value)
{
    if (int rc = check_fixed(t, f, order, low, up, bounds_stride, params, params_stride, batch, out_value)) return rc;
    HostCtx* hc = nullptr;
    if (int rc = host_ctx(&hc)) return rc;
    cudaStream_t g_stream = hc->stream;  // this thread's stream
    const size_t es = dtype_size(t->dtype);
    const int dim = kind_dim(f->kind);
    const size_t nb = bounds_stride ? (size_t)batch * bounds_stride : (size_t)dim;
    const size_t np = f->nparams ? (params_stride ? (size_t)batch * params_stride : (size_t)f->nparams) : 0;
    // per-integral operands and the results in pinned host memory are used in place (see
    // fts_adaptive_batch); anything else is staged through device scratch
    void *dl = nullptr, *du = nullptr, *dp = nullptr, *dv = nullptr;
    if (batch > 0 && bounds_stride && input_in_place()) { dl = mapped_dev_ptr(low); du = mapped_dev_ptr(up); }
    if (batch > 0 && np && params_stride && input_in_place()) dp = mapped_dev_ptr(params);
    if (batch > 0 && output_in_place()) dv = mapped_dev_ptr(out_value);
    const bool out_in_place = dv != nullptr;
    if (!dl) if (int rc = stage_in(hc, S_LOW, low, nb * es, &dl)) return rc;
    if (!du) if (int rc = stage_in(hc, S_UP, up, nb * es, &du)) return rc;
    if (!dp) if (int rc = stage_in(hc, S_PAR, params, np * es, &dp)) return rc;
    if (!dv) if (int rc = stage_out(hc, S_VAL, true, (size_t)batch * es, &dv)) return rc;
    if (int rc = fts_integrate_batch_dev(t, f, order, dl, du, bounds_stride, dp, params_stride, batch, dv, g_stream)) return rc;
    if (batch && !out_in_place) FTS_CUDA(cudaMemcpyAsync(out_value, dv, (size_t)batch * es, cudaMemcpyDeviceToHost, g_stream));
    FTS_CUDA(cudaStreamSynchronize(g_stream));
    return FTS_OK;
}

// integrate1D_cmpl(T, f, N) (integrate1D.jl:49-65): f(x, 1-x, 1+x) over [-1, 1] on the fixed grid, one
// integral per parameter row
template <class T> static int run_fixed_cmpl(const void* fn, fts_tables t, const void* params, int pstride, long long batch, void* out, cudaStream_t s)
{
    fts::Args<T> a;
    memset(&a, 0, sizeof a);
    a.grid = (const fts::Node<T>*)t->d_grid;
    a.n = t->n;
    a.h = load_elem<T>(t->h);
    a.params = (const T*)params; a.pstride = pstride; a.batch = batch; a.value = (T*)out;
    return launch(fn, batch, 256, &a, s);
}

static int check_fixed_cmpl(fts_tables t, fts_integrand f, int order, const void* params, int pstride, long long batch, void* out)
{
    if (!t || !f) return set_error(FTS_ERR_INVALID_ARGUMENT, "null handle");
    if (f->kind != FTS_KIND_1D_CMPL) return set_error(FTS_ERR_DIMENSION_MISMATCH, "integrate1D_cmpl needs an integrand of kind f(x, bmx, xma)");
    if (t->dtype == FTS_DD64) return set_error(FTS_ERR_UNSUPPORTED, "integrate1D_cmpl: f32/f64");
    if (f->nparams > 0 && (!params || (pstride != 0 && pstride < f->nparams))) return set_error(FTS_ERR_INVALID_ARGUMENT, "integrand needs " + std::to_string(f->nparams) + " parameters per integral");
    if (batch < 0 || (batch > 0 && !out)) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad batch arguments");
    if (order != FTS_ORDER_REFERENCE && order != FTS_ORDER_FAST && order != FTS_ORDER_AUTO) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad order");
    return ensure_init();
}

extern "C" int fts_integrate1d_cmpl_batch_dev(fts_tables t, fts_integrand f, int order, const void* params, int params_stride, int64_t batch,
                                              void* out_value, void* stream)
{
    if (int rc = check_fixed_cmpl(t, f, order, params, params_stride, batch, out_value)) return rc;
    if (batch == 0) return FTS_OK;
    const void* fn = integrand_kernel(f, t->dtype, KID_FIXED_CMPL, order == FTS_ORDER_FAST ? 1 : 0);
    if (!fn) return set_error(FTS_ERR_UNSUPPORTED, "no integrate1D_cmpl kernel for this integrand/dtype");
    cudaStream_t s = (cudaStream_t)stream;
    return t->dtype == FTS_F32 ? run_fixed_cmpl<float>(fn, t, params, params_stride, batch, out_value, s)
                               : run_fixed_cmpl<double>(fn, t, params, params_stride, batch, out_value, s);
}

extern "C" int fts_integrate1d_cmpl_batch(fts_tables t, fts_integrand f, int order, const void* params, int params_stride, int64_t batch, void* out_value)
{
    if (int rc = check_fixed_cmpl(t, f, order, params, params_stride, batch, out_value)) return rc;
    HostCtx* hc = nullptr;
    if (int rc = host_ctx(&hc)) return rc;
    const size_t es = dtype_size(t->dtype);
    const size_t np = f->nparams ? (params_stride ? (size_t)batch * params_stride : (size_t)f->nparams) : 0;
    void *dp = nullptr, *dv = nullptr;
    if (int rc = stage_in(hc, S_PAR, params, np * es, &dp)) return rc;
    if (int rc = stage_out(hc, S_VAL, true, (size_t)batch * es, &dv)) return rc;
    if (int rc = fts_integrate1d_cmpl_batch_dev(t, f, order, dp, params_stride, batch, dv, hc->stream)) return rc;
    if (batch) FTS_CUDA(cudaMemcpyAsync(out_value, dv, (size_t)batch * es, cudaMemcpyDeviceToHost, hc->stream));
    FTS_CUDA(cudaStreamSynchronize(hc->stream));
    return FTS_OK;
}

// ---- whole-grid level machinery (one large 2D/3D integral; also the multi-GPU shard path) ----
template <class T, int D>
__global__ void k_state_to_result(const T* state, const T* low, const T* up, int max_levels, int options, T* value, T* err, int* levels, int* flags)
{
    const int done = (int)fts::num<T>::to_double(state[5]);
    int fl = done == 2 ? (fts::FLAG_ZERO_WIDTH | fts::FLAG_CONVERGED) : (done == 1 ? fts::FLAG_CONVERGED : (max_levels > 0 ? fts::FLAG_MAX_LEVELS : 0));
    if (options & fts::OPT_QUAD_EDGE) {  // flipped bounds give the negated value through the signed half-widths
        bool neg = false;
        for (int d = 0; d < D; ++d) neg = neg != (low[d] > up[d]);
        if (neg) fl |= fts::FLAG_FLIPPED;
    }
    *value = state[3];
    if (err) *err = state[4];
    if (levels) *levels = (int)fts::num<T>::to_double(state[6]);
    if (flags) *flags = fl;
}

static size_t grid_level_smem(int dim, bool cmpl, int level, size_t es)
{
    const size_t n = level == 0 ? 2 : ((size_t)2 << level);
    return (dim == 3 ? fts::SM3_ARRAYS : (cmpl ? 9 : 5)) * n * es;
}

static unsigned grid_level_blocks(int dim, int level, int nranks)
{
    const double n = level == 0 ? 2.0 : (double)(2 << level);
    const double cells = (dim == 3 ? 0.875 * n * n * n + 0.75 * n * n : 0.75 * n * n + n) / nranks;
    double want = cells / (fts::CTA_THREADS * 2.0);  // >= 2 cells per thread
    const double cap = (double)g_prop.multiProcessorCount * 4;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (unsigned)want;
}

template <class T>
static int launch_level(const void* fn, fts_cache c, int dim, bool cmpl, const T* low, const T* up, const T* params, int level, int rank,
                        int nranks, T* block_partials, unsigned int* ticket, T* partial_out, const T* state, cudaStream_t s)
{
    fts::Args<T> a;
    memset(&a, 0, sizeof a);
    fill_cache_args(a, c);
    a.low = low; a.up = up; a.params = params; a.bstride = dim; a.pstride = 0; a.batch = 1;
    a.level = level; a.rank = rank; a.nranks = nranks;
    a.block_partials = block_partials; a.ticket = ticket; a.partial_out = partial_out; a.state = state;
    return launch_raw(fn, grid_level_blocks(dim, level, nranks), fts::CTA_THREADS, grid_level_smem(dim, cmpl, level, sizeof(T)), &a, s);
}

// Levels lo ... hi of one integral in ONE launch of the level kernel (stop test in its last CTA).  More than one
// CTA: cooperative launch + the generation barrier `gen`.  FTS_B200_GRID_RANGE = the deepest level that joins
// the first launch (default 7: deeper levels last 100 us and more, they get a launch and a full-size grid each).
static int range_top_level()
{
    static const int v = [] { const char* e = getenv("FTS_B200_GRID_RANGE"); const int x = e ? atoi(e) : 7; return x < 0 ? 0 : (x > 30 ? 30 : x); }();
    return v;
}
template <class T>
static int launch_level_range(const void* fn, fts::Args<T>& a, int dim, bool cmpl, int lo, int hi, unsigned int* gen, cudaStream_t s)
{
    unsigned blocks = 1;
    for (int l = lo; l <= hi; ++l) {
        const unsigned b = grid_level_blocks(dim, l, l >= a.shard_from ? a.nranks : 1);
        if (b > blocks) blocks = b;
        a.range_blocks[l & 31] = (unsigned short)b;
    }
    a.level = lo; a.level_hi = hi; a.step_fused = 1;
    const size_t smem = grid_level_smem(dim, cmpl, hi, sizeof(T));
    if (blocks == 1 || lo == hi) {
        a.gen = nullptr;
        return launch_raw(fn, blocks, fts::CTA_THREADS, smem, &a, s);
    }
    a.gen = gen;
    return launch_cooperative(fn, blocks, fts::CTA_THREADS, smem, &a, s);
}

// FTS_B200_STEP_FUSED=0: the two-launch form (level kernel, then a one-warp stop-test kernel) for A/B runs
static bool step_in_level_kernel()
{
    static const bool v = [] { const char* e = getenv("FTS_B200_STEP_FUSED"); return !(e && atoi(e) == 0); }();
    return v;
}

template <class T, int D> static void launch_step(const T* low, const T* up, T tm, T rtol, T atol, int level, int nranks, const T* partials, T* state, cudaStream_t s, int options = 0, int max_levels = 0)
{
    fts::k_sharded_step<T, D><<<1, 1, 0, s>>>(low, up, tm, rtol, atol, level, nranks, partials, state, options, max_levels);
    g_launches.fetch_add(1);
}

// adaptive integration of a SMALL batch of large 2D/3D integrals: per integral, one whole-grid
// launch per level + a one-thread stop test, all queued on the stream without host round trips
// (later levels early-exit once the device-side `done` flag is set).
template <class T>
static int run_adaptive_grid(const void* fn, fts_cache c, fts_integrand f, int options, const T* low, const T* up, int bstride,
                             const T* params, int pstride, T rtol, T atol, int max_levels, long long batch, T* out_value, T* out_err,
                             int32_t* out_levels, int32_t* out_flags, cudaStream_t s)
{
    const int dim = kind_dim(f->kind);
    const bool cmpl = kind_cmpl(f->kind);
    const size_t nblk = (size_t)g_prop.multiProcessorCount * 4;
    T* scratch = nullptr;
    FTS_CUDA(cudaMallocAsync((void**)&scratch, (2 * nblk + 2 + 8 + 2) * sizeof(T) + 256, s));
    T* block_partials = scratch;
    T* partial = scratch + 2 * nblk;
    T* state = partial + 2;
    unsigned int* ticket = (unsigned int*)(state + 8);  // ticket, and 128 B behind it (its own line: polled by every CTA) the generation barrier
    unsigned int* gen = ticket + 32;
    FTS_CUDA(cudaMemsetAsync(ticket, 0, 33 * sizeof(unsigned int), s));
    const T tm = load_elem<T>(c->tm);
    const bool one_launch = step_in_level_kernel();
    for (long long b = 0; b < batch; ++b) {
        const T* lo = low + b * bstride;
        const T* hi = up + b * bstride;
        const T* pp = params ? params + b * pstride : nullptr;
        for (int level = 0; level <= max_levels; ++level) {
            if (one_launch) {  // stop test in the last CTA of the level kernel, levels 0 ... range_top_level() in one launch
                fts::Args<T> a;
                memset(&a, 0, sizeof a);
                fill_cache_args(a, c);
                a.low = lo; a.up = hi; a.params = pp; a.bstride = dim; a.pstride = 0; a.batch = 1;
                a.rank = 0; a.nranks = 1;
                a.block_partials = block_partials; a.ticket = ticket; a.partial_out = partial; a.state = level ? state : nullptr;
                int top = level;
                if (level == 0) top = max_levels < range_top_level() ? max_levels : range_top_level();
                a.state_out = state;
                a.gen_base = (unsigned int)b * 64u;
                a.rtol = rtol; a.atol = atol; a.options = options; a.max_levels = max_levels;
                if (int rc = launch_level_range<T>(fn, a, dim, cmpl, level, top, gen, s)) return rc;
                level = top;
                continue;
            }
            if (int rc = launch_level<T>(fn, c, dim, cmpl, lo, hi, pp, level, 0, 1, block_partials, ticket, partial, level ? state : nullptr, s)) return rc;
            if (dim == 3) launch_step<T, 3>(lo, hi, tm, rtol, atol, level, 1, partial, state, s, options, max_levels);
            else launch_step<T, 2>(lo, hi, tm, rtol, atol, level, 1, partial, state, s, options, max_levels);
        }
        if (dim == 3) k_state_to_result<T, 3><<<1, 1, 0, s>>>(state, lo, hi, max_levels, options, out_value + b, out_err ? out_err + b : nullptr,
                                                              out_levels ? out_levels + b : nullptr, out_flags ? out_flags + b : nullptr);
        else k_state_to_result<T, 2><<<1, 1, 0, s>>>(state, lo, hi, max_levels, options, out_value + b, out_err ? out_err + b : nullptr,
                                                     out_levels ? out_levels + b : nullptr, out_flags ? out_flags + b : nullptr);
        g_launches.fetch_add(1);
    }
    FTS_CUDA(cudaGetLastError());
    FTS_CUDA(cudaFreeAsync(scratch, s));
    return FTS_OK;
}

// Persistent per-stream queues of the deep-level hand-off and of the bucket passes:
// {count, ticket, sort_ctl[2], pad[4], idx[capacity], bucket arrays[2 * capacity + capacity / 16] (see Args::sort_idx)}.  Launches on one stream are
// ordered, so they can share the queues (the tail kernel re-arms them when it leaves); different streams
// get different ones.
constexpr int QUEUE_HEADER = 8;
struct TailQueue {
    int* buf = nullptr;
    size_t cap = 0;
    int* fb_host = nullptr;  // pinned, device-mapped word: did the last call on this stream see incoherent warps? (see run_adaptive)
    int* fb_dev = nullptr;
};
static std::mutex g_tail_mu;
static std::vector<std::pair<cudaStream_t, TailQueue>> g_tail_queues;
static int tail_queue_for(cudaStream_t s, size_t batch, int** out, size_t* cap_out = nullptr, TailQueue* q_out = nullptr)
{
    std::lock_guard<std::mutex> lk(g_tail_mu);
    TailQueue* q = nullptr;
    for (auto& e : g_tail_queues)
        if (e.first == s) q = &e.second;
    if (!q) {
        g_tail_queues.emplace_back(s, TailQueue());
        q = &g_tail_queues.back().second;
    }
    if (q->cap < batch) {
        if (q->buf) {
            FTS_CUDA(cudaStreamSynchronize(s));
            FTS_CUDA(cudaFree(q->buf));
            q->buf = nullptr;
            q->cap = 0;
        }
        const size_t cap = (batch + batch / 4 + 2 * 1024) / 1024 * 1024;
        FTS_CUDA(cudaMalloc((void**)&q->buf, sizeof(int) * (3 * cap + cap / 16 + QUEUE_HEADER)));
        FTS_CUDA(cudaMemsetAsync(q->buf, 0, sizeof(int) * QUEUE_HEADER, s));  // ordered before the first kernel on `s` (non-blocking streams do not wait for the legacy stream)
        q->cap = cap;
    }
    if (!q->fb_host) {
        FTS_CUDA(cudaHostAlloc((void**)&q->fb_host, sizeof(int), cudaHostAllocMapped));
        *q->fb_host = 0;
        FTS_CUDA(cudaHostGetDevicePointer((void**)&q->fb_dev, q->fb_host, 0));
    }
    *out = q->buf;
    if (cap_out) *cap_out = q->cap;
    if (q_out) *q_out = *q;  // a copy: the vector may grow under another thread once the lock is gone
    return FTS_OK;
}

// fts_shutdown / re-binding to another device: the queues are keyed by raw stream pointers, which a
// later stream (possibly on another device) may reuse, and their buffers belong to the old device
static void drop_device_state()
{
    {
        std::lock_guard<std::mutex> lk(g_tail_mu);
        for (auto& e : g_tail_queues) {
            if (e.second.buf) cudaFree(e.second.buf);
            if (e.second.fb_host) cudaFreeHost(e.second.fb_host);
        }
        g_tail_queues.clear();
    }
    if (g_peer_wait) cudaFree(g_peer_wait);
    g_peer_wait = nullptr;
    cudaGetLastError();
}

// kernel style an adaptive launch will use (shared by run_adaptive and the host-buffer entry point)
static Style adaptive_style(fts_cache c, fts_integrand f, int order, int max_levels, long long batch, const void** fn_cta_out,
                            const void** fn_grid_out, size_t* smem_cta_out)
{
    const int dim = kind_dim(f->kind);
    const bool cmpl = kind_cmpl(f->kind);
    const bool is_dd = c->dtype == FTS_DD64;
    const size_t es = dtype_size(c->dtype);
    const void* fn_cta = is_dd ? nullptr : integrand_kernel(f, c->dtype, KID_ADAPT_CTA, 1);
    const void* fn_grid = (is_dd || dim == 1) ? nullptr : integrand_kernel(f, c->dtype, KID_LEVEL_GRID, 1);
    const size_t nmax = max_levels > 0 ? ((size_t)2 << max_levels) : 2;
    // 2D: derived coordinate arrays + two raw node slices + two mbarriers (fts::sm2_group_bytes)
    const size_t node_b = (cmpl ? 4 : 2) * es;
    const size_t smem_cta = dim == 1 ? 0 : (dim == 3 ? fts::SM3_ARRAYS * nmax * es : (cmpl ? 9 : 5) * nmax * es + 2 * nmax * node_b + 16);
    const bool cta_ok = fn_cta && smem_cta <= dyn_smem_limit();
    const bool grid_ok = fn_grid && grid_level_smem(dim, cmpl, max_levels, es) <= dyn_smem_limit();
    if (fn_cta_out) *fn_cta_out = fn_cta;
    if (fn_grid_out) *fn_grid_out = fn_grid;
    if (smem_cta_out) *smem_cta_out = smem_cta;
    return pick_style(order, dim, batch, max_levels, cta_ok, grid_ok);
}

template <class T>
static int run_adaptive(fts_cache c, fts_integrand f, int order, int options, const void* low, const void* up, int bstride,
                        const void* params, int pstride, const void* rtol, const void* atol, int max_levels, long long batch,
                        void* out_value, void* out_err, int32_t* out_levels, int32_t* out_flags, cudaStream_t s)
{
    if (batch == 0) return FTS_OK;
    const int dim = kind_dim(f->kind);
    const bool cmpl = kind_cmpl(f->kind);
    const bool is_dd = c->dtype == FTS_DD64;
    const void *fn_cta = nullptr, *fn_grid = nullptr;
    size_t smem_cta = 0;
    const Style st = adaptive_style(c, f, order, max_levels, batch, &fn_cta, &fn_grid, &smem_cta);
    if (st == STYLE_GRID)
        return run_adaptive_grid<T>(fn_grid, c, f, options, (const T*)low, (const T*)up, bstride, (const T*)params, pstride, load_elem<T>(rtol),
                                    load_elem<T>(atol), max_levels, batch, (T*)out_value, (T*)out_err, out_levels, out_flags, s);
    fts::Args<T> a;
    memset(&a, 0, sizeof a);
    fill_cache_args(a, c);
    a.low = (const T*)low; a.up = (const T*)up; a.bstride = bstride;
    a.params = (const T*)params; a.pstride = pstride;
    a.rtol = load_elem<T>(rtol); a.atol = load_elem<T>(atol);
    a.max_levels = max_levels; a.options = options; a.batch = batch;
    a.value = (T*)out_value; a.err = (T*)out_err; a.levels = out_levels; a.flags = out_flags;
    if (st == STYLE_CTA) {
        if (batch > 0x7fffffffLL) return set_error(FTS_ERR_INVALID_ARGUMENT, "batch too large for one launch");
        // 2D batches large enough to fill the GPU with warps: one WARP per integral up to level
        // `lw` (no CTA barriers; shallow integrals such as config C stop at level 4), the rest is
        // queued and finished by whole CTAs in a second launch of the same kernel
        // default 4: 4.4 KiB of shared memory per warp, seven 4-warp CTAs per SM — 4096 integrals are ONE wave
        static const int kWarpLevels = [] { const char* e = getenv("FTS_B200_WARP2D_LEVELS"); const int v = e ? atoi(e) : 4; return v < 0 ? -1 : (v > 7 ? 7 : v); }();
        const void* fn_warp = (dim == 2 && kWarpLevels >= 0 && batch >= 8LL * g_prop.multiProcessorCount) ? integrand_kernel(f, c->dtype, KID_ADAPT_WARP, 1) : nullptr;
        if (dim == 2 && kWarpLevels >= 0 && !fn_warp) g_last_error.clear();
        if (fn_warp) {
            const int lw = max_levels < kWarpLevels ? max_levels : kWarpLevels;
            const int warps = fts::WARP2D_WARPS;
            const size_t ncap_w = (size_t)2 << lw;
            // per warp (fts::warp2d_bytes): x-side records + y-side arrays of ncap + 1 entries, two raw node slices, two mbarriers
            const size_t nc_w = ncap_w + 1, rec_w = sizeof(T) == 8 ? 6 : 8;
            const size_t xy_w = nc_w * rec_w * sizeof(T) + ((5 * nc_w * sizeof(T) + 15) & ~(size_t)15), fb_w = (cmpl ? 9 : 5) * ncap_w * sizeof(T);
            const size_t smem_w = (size_t)warps * ((xy_w > fb_w ? xy_w : fb_w) + 2 * ncap_w * (cmpl ? 4 : 2) * sizeof(T) + 16);
            int* queue = nullptr;
            if (int rc = tail_queue_for(s, (size_t)batch, &queue)) return rc;
            a.tail_count = queue;
            a.ticket = (unsigned int*)(queue + 1);
            a.tail_idx = queue + QUEUE_HEADER;
            a.group = 32;
            a.level = lw;
            // a programmatic dependent of the previous kernel on the stream (it waits in `griddepcontrol.wait` before it
            // touches global memory): in a loop of calls the set-up of this grid hides behind the previous call's queue pass
            if (int rc = launch_dependent(fn_warp, (unsigned)((batch + warps - 1) / warps), warps * 32, smem_w, &a, s)) return rc;
            if (lw == max_levels) return FTS_OK;  // nothing can be queued
            // second pass: whole CTAs walk the queue of deeper integrals.  Two CTAs per SM: with an empty queue
            // (every integral converged in the warp pass) the launch is over in a couple of microseconds
            a.group = 0;
            a.level = 0;
            const long long cap = (long long)g_prop.multiProcessorCount * 2;
            return launch_dependent(fn_cta, (unsigned)(batch < cap ? batch : cap), fts::CTA_THREADS, smem_cta, &a, s);
        }
        return launch_raw(fn_cta, (unsigned)batch, fts::CTA_THREADS, smem_cta, &a, s);
    }
    // REFERENCE order, 2D/3D, too few integrals to fill the GPU with one thread each: one CTA per
    // integral evaluates the level's terms in parallel and adds them in the reference's order
    // (bit-identical to the thread-per-integral kernel, see k_adaptive2d_ord)
    if (order == FTS_ORDER_REFERENCE && dim >= 2 && !is_dd && batch < 64LL * g_prop.multiProcessorCount) {
        static const bool use_ord = [] { const char* e = getenv("FTS_B200_ORDERED"); return !e || atoi(e) != 0; }();
        const void* fn_ord = use_ord ? integrand_kernel(f, c->dtype, KID_ADAPT_ORD, 0) : nullptr;
        if (fn_ord) {
            // clusters of 8/4/2 CTAs when the integrals are too few to occupy the SMs one CTA each
            static const int max_cluster = [] { const char* e = getenv("FTS_B200_ORD_CLUSTER"); const int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
            unsigned csize = 1;
            while ((int)csize * 2 <= max_cluster && batch * (long long)csize * 2 <= 2LL * g_prop.multiProcessorCount) csize *= 2;
            // tile buffers: 2 x 4096 terms for a few integrals (the evaluation of a whole tile overlaps
            // the add of the previous one), 2 x 1024 when there are CTAs to spare (16 KiB: more CTAs per SM)
            a.level = batch >= g_prop.multiProcessorCount ? 1024 : 4096;
            return launch_cluster(fn_ord, (unsigned)batch, csize, 256, 2 * (size_t)a.level * sizeof(T), &a, s);
        }
        if (use_ord) g_last_error.clear();  // families without the kernel fall back to thread per integral
    }
    const int fast = order == FTS_ORDER_FAST ? 1 : 0;
    const void* fn = integrand_kernel(f, c->dtype, KID_ADAPT_TPI, fast);
    if (!fn) return set_error(FTS_ERR_UNSUPPORTED, "no adaptive kernel built for this integrand/dtype/order");
    // 1D: integrals still unconverged after `tail_level` are finished by one CTA each
    // (k_adaptive1d_tail, same summation order) instead of stalling their warp for 2^level steps
    const int tail_level = is_dd ? 7 : 8;
    const void* fn_tail = (dim == 1 && max_levels > tail_level) ? integrand_kernel(f, c->dtype, KID_ADAPT_TAIL, fast) : nullptr;
    a.tail_level = -1;
    const void* fn_bucket = nullptr;
    const void* fn_sort_all = nullptr;  // set: the whole batch goes through the bucket passes, k_adaptive1d_tpi is not launched
    if (fn_tail) {
        if (batch > 0x7fffffffLL) return set_error(FTS_ERR_INVALID_ARGUMENT, "batch too large for one launch");
        int* queue = nullptr;
        size_t qcap = 0;
        TailQueue tq;
        if (int rc = tail_queue_for(s, (size_t)batch, &queue, &qcap, &tq)) return rc;
        a.tail_level = tail_level;
        a.tail_count = queue;
        a.ticket = (unsigned int*)(queue + 1);
        a.sort_ctl = queue + 2;
        a.tail_idx = queue + QUEUE_HEADER;
        // Bucket passes (Float64 exp families, big batches): warps whose lanes hold unrelated integrals leave them to
        // k_bucket_sort + k_adaptive1d_bucket, which regroup them.  Two launches that a batch in coherent order does not
        // need (5 us of a 280 us call when they find nothing to do), so they are armed by feedback: every call detects
        // incoherent warps (k_adaptive1d_tpi) and the tail kernel leaves the finding in a pinned word; a call arms the
        // passes when the last finished call on its stream found such warps.  Results do not depend on the choice, bit
        // for bit.  The finding is a (sampled) COUNT of incoherent warps: when they were the majority, the whole batch
        // goes through the passes (k_bucket_sort_all computes the keys itself) and k_adaptive1d_tpi — which would only
        // read the parameters and write the keys, at the price of 8192 CTA launches per 2^20 integrals — is not launched.
        // FTS_B200_BUCKET = 0: never armed, 2: always armed, 3: always the whole batch, 1 (default): by feedback.
        static const int bucket_mode = [] { const char* e = getenv("FTS_B200_BUCKET"); return e ? atoi(e) : 1; }();
        if (bucket_mode != 0 && !cmpl && c->dtype == FTS_F64 && batch >= 4LL * g_prop.multiProcessorCount * 256) {
            fn_bucket = integrand_kernel(f, c->dtype, KID_ADAPT_BUCKET, fast);
            if (!fn_bucket) g_last_error.clear();
            else {
                // the two kernels are loaded (lazy module loading: ~0.4 ms) while the passes are not armed yet, not
                // in the first call that needs them
                static std::atomic<const void*> loaded{nullptr};
                if (loaded.load() != fn_bucket) {
                    cudaFuncAttributes fa;
                    cudaFuncGetAttributes(&fa, (const void*)&fts::k_bucket_sort<void>);
                    cudaFuncGetAttributes(&fa, fn_bucket);
                    if (const void* fsa = integrand_kernel(f, c->dtype, KID_BUCKET_SORT_ALL, 0)) cudaFuncGetAttributes(&fa, fsa);
                    else g_last_error.clear();
                    cudaGetLastError();
                    loaded.store(fn_bucket);
                }
                a.sort_fb = tq.fb_dev;
                const long long found = *(volatile int*)tq.fb_host;  // ~ incoherent warps of the last finished call on this stream
                if (bucket_mode >= 2 || found != 0) {
                    a.sort_idx = queue + QUEUE_HEADER + qcap;
                    if (bucket_mode == 3 || (bucket_mode == 1 && found * 64 >= batch)) {
                        fn_sort_all = integrand_kernel(f, c->dtype, KID_BUCKET_SORT_ALL, 0);
                        if (!fn_sort_all) g_last_error.clear();
                    }
                } else {
                    fn_bucket = nullptr;  // detect only
                }
            }
        }
    }
    // Double64, 1D: a group of 4 (FTS_B200_DD_GROUP = 2: 2, 1: off) lanes per integral while that still leaves the
    // GPU under a few waves (k_adaptive1d_grp)
    long long threads = batch;
    if (is_dd && dim == 1) {
        static const int dd_group = [] { const char* e = getenv("FTS_B200_DD_GROUP"); const int v = e ? atoi(e) : 4; return v == 2 || v == 4 ? v : 1; }();
        if (dd_group > 1 && batch * dd_group <= (1LL << 21)) {
            const void* fn_grp = integrand_kernel(f, c->dtype, KID_ADAPT_GRP, dd_group == 2 ? 1 : 0);
            if (fn_grp) {
                fn = fn_grp;
                threads = batch * dd_group;
            } else {
                g_last_error.clear();
            }
        }
    }
    const unsigned chunks = (unsigned)((batch + fts::SORT_CHUNK - 1) / fts::SORT_CHUNK);
    if (fn_sort_all) {
        if (int rc = launch_dependent(fn_sort_all, chunks, 256, 0, &a, s)) return rc;
        if (int rc = launch_dependent(fn_bucket, (unsigned)g_prop.multiProcessorCount * 2, 256, 0, &a, s)) return rc;
    } else {
        if (int rc = launch(fn, threads, tpi_block(dim, c->dtype), &a, s)) return rc;
        if (fn_bucket) {
            const int* ctl = a.sort_ctl;
            int* base = a.sort_idx;
            long long nb = batch;
            void* params[] = {(void*)&ctl, (void*)&base, (void*)&nb};
            if (int rc = launch_dependent_params((const void*)&fts::k_bucket_sort<void>, chunks, 256, params, s)) return rc;
            if (int rc = launch_dependent(fn_bucket, (unsigned)g_prop.multiProcessorCount * 2, 256, 0, &a, s)) return rc;
        }
    }
    if (fn_tail) {
        const long long cap = (long long)g_prop.multiProcessorCount * 8;
        if (int rc = launch_dependent(fn_tail, (unsigned)(batch < cap ? batch : cap), 256, 0, &a, s)) return rc;
    }
    return FTS_OK;
}

static int check_cache_match(fts_cache c, fts_integrand f)
{
    const int dim = kind_dim(f->kind);
    const int cdim = c->D == 0 ? 1 : c->D;
    if (dim != cdim) return set_error(FTS_ERR_DIMENSION_MISMATCH, "integrand is " + std::to_string(dim) + "D but the cache is " + std::to_string(cdim) + "D");
    if (kind_cmpl(f->kind) && c->kind != 1) return set_error(FTS_ERR_INVALID_ARGUMENT, "complement integrands need a complement cache (adaptive_cache_1D(T; complement=true))");
    return FTS_OK;
}

static int check_adaptive(fts_cache c, fts_integrand f, int order, const void* low, const void* up, int bstride, const void* params,
                          int pstride, const void* rtol, const void* atol, int max_levels, long long batch, void* out_value)
{
    if (!c || !f) return set_error(FTS_ERR_INVALID_ARGUMENT, "null handle");
    if (int rc = check_cache_match(c, f)) return rc;
    const int dim = kind_dim(f->kind);
    if (max_levels < 0) return set_error(FTS_ERR_INVALID_ARGUMENT, "`max_levels` must be nonnegative.");
    if (c->max_levels < max_levels)  // caches.jl:25-29
        return set_error(FTS_ERR_INVALID_ARGUMENT, "provided adaptive cache supports " + std::to_string(c->max_levels) + " levels, but `max_levels=" + std::to_string(max_levels) + "` was requested.");
    if (!rtol || !atol) return set_error(FTS_ERR_INVALID_ARGUMENT, "rtol/atol pointers are required");
    if (!(elem_hi(c->dtype, rtol) >= 0.0)) return set_error(FTS_ERR_INVALID_ARGUMENT, "`rtol` must be nonnegative.");  // core.jl:41
    if (!(elem_hi(c->dtype, atol) >= 0.0)) return set_error(FTS_ERR_INVALID_ARGUMENT, "`atol` must be nonnegative.");  // core.jl:42
    if (bstride != 0 && bstride != dim) return set_error(FTS_ERR_DIMENSION_MISMATCH, "low/up must have length " + std::to_string(dim) + " per integral");
    if (f->nparams > 0 && (!params || (pstride != 0 && pstride < f->nparams))) return set_error(FTS_ERR_INVALID_ARGUMENT, "integrand needs " + std::to_string(f->nparams) + " parameters per integral");
    if (batch < 0 || !low || !up || (batch > 0 && !out_value)) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad batch arguments");
    if (order != FTS_ORDER_REFERENCE && order != FTS_ORDER_FAST && order != FTS_ORDER_AUTO) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad order");
    return ensure_init();
}

extern "C" int fts_adaptive_batch_dev(fts_cache c, fts_integrand f, int order, int options, const void* low, const void* up,
                                      int bounds_stride, const void* params, int params_stride, const void* rtol, const void* atol,
                                      int max_levels, int64_t batch, void* out_value, void* out_err, int32_t* out_levels,
                                      int32_t* out_flags, void* stream)
{
    if (int rc = check_adaptive(c, f, order, low, up, bounds_stride, params, params_stride, rtol, atol, max_levels, batch, out_value)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    switch (c->dtype) {
    case FTS_F32: return run_adaptive<float>(c, f, order, options, low, up, bounds_stride, params, params_stride, rtol, atol, max_levels, batch, out_value, out_err, out_levels, out_flags, s);
    case FTS_F64: return run_adaptive<double>(c, f, order, options, low, up, bounds_stride, params, params_stride, rtol, atol, max_levels, batch, out_value, out_err, out_levels, out_flags, s);
    case FTS_DD64: return run_adaptive<fts::dd>(c, f, order, options, low, up, bounds_stride, params, params_stride, rtol, atol, max_levels, batch, out_value, out_err, out_levels, out_flags, s);
    }
    return set_error(FTS_ERR_INTERNAL, "bad dtype");
}

// Host-buffer entry point (pinned buffers: fts_host_alloc / cudaHostRegister; pageable buffers work
// too, their copies then stage through the driver).
extern "C" int fts_adaptive_batch(fts_cache c, fts_integrand f, int order, int options, const void* low, const void* up,
                                  int bounds_stride, const void* params, int params_stride, const void* rtol, const void* atol,
                                  int max_levels, int64_t batch, void* out_value, void* out_err, int32_t* out_levels, int32_t* out_flags)
{
    // FTS_B200_TRACE=1: host-side phases of this entry point (microseconds, averaged over 1000 calls, to stderr)
    static const bool trace = [] { const char* e = getenv("FTS_B200_TRACE"); return e && atoi(e) != 0; }();
    thread_local double tr_acc[4] = {0, 0, 0, 0};
    thread_local int tr_n = 0;
    const auto tr0 = std::chrono::steady_clock::now();
    auto tr_mark = [&](int k, std::chrono::steady_clock::time_point& last) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        tr_acc[k] += std::chrono::duration<double, std::micro>(now - last).count();
        last = now;
    };
    auto tr_last = tr0;
    if (int rc = check_adaptive(c, f, order, low, up, bounds_stride, params, params_stride, rtol, atol, max_levels, batch, out_value)) return rc;
    HostCtx* hc = nullptr;
    if (int rc = host_ctx(&hc)) return rc;
    cudaStream_t g_stream = hc->stream;  // this thread's stream, pipeline streams and scratch
    cudaStream_t* g_pipe = hc->pipe;
    cudaEvent_t g_pipe_ev = hc->ev;
    Scratch* g_scratch = hc->scratch;
    const size_t es = dtype_size(c->dtype);
    const int dim = kind_dim(f->kind);
    const size_t nb = bounds_stride ? (size_t)batch * bounds_stride : (size_t)dim;
    const size_t np = f->nparams ? (params_stride ? (size_t)batch * params_stride : (size_t)f->nparams) : 0;

    // Every operand is either used IN PLACE (pinned, mapped host memory) or STAGED through device
    // scratch — see zero_copy_mode.  Broadcast operands (a few bytes every thread reads) are
    // always staged.
    struct Operand {
        const void* host; char* dev; size_t elem; bool staged;
    };
    auto resolve = [&](int slot, const void* host, size_t bytes, size_t elem, bool per_integral, bool is_output, Operand& o) -> cudaError_t {
        o = {host, nullptr, elem, false};
        if (!host || bytes == 0) return cudaSuccess;
        if (batch > 0 && per_integral && (is_output ? output_in_place() : input_in_place())) o.dev = (char*)mapped_dev_ptr(host);
        if (o.dev) return cudaSuccess;
        o.staged = true;
        cudaError_t e = g_scratch[slot].need(bytes);
        o.dev = (char*)g_scratch[slot].p;
        return e;
    };
    Operand ol, ou, op_, ov, oe, olv, ofl;
    FTS_CUDA(resolve(S_LOW, low, nb * es, (size_t)bounds_stride * es, bounds_stride != 0, false, ol));
    FTS_CUDA(resolve(S_UP, up, nb * es, (size_t)bounds_stride * es, bounds_stride != 0, false, ou));
    FTS_CUDA(resolve(S_PAR, np ? params : nullptr, np * es, (size_t)params_stride * es, params_stride != 0, false, op_));
    FTS_CUDA(resolve(S_VAL, out_value, (size_t)batch * es, es, true, true, ov));
    FTS_CUDA(resolve(S_ERR, out_err, (size_t)batch * es, es, true, true, oe));
    FTS_CUDA(resolve(S_LVL, out_levels, (size_t)batch * 4, 4, true, true, olv));
    FTS_CUDA(resolve(S_FLG, out_flags, (size_t)batch * 4, 4, true, true, ofl));
    tr_mark(0, tr_last);  // checks + operand resolution
    Operand* ins[3] = {&ol, &ou, &op_};
    Operand* outs[4] = {&ov, &oe, &olv, &ofl};
    bool staged_in = false, staged_out = false;  // staged per-integral operands -> chunk pipeline
    for (Operand* o : ins) staged_in |= o->staged && o->elem != 0;
    for (Operand* o : outs) staged_out |= o->staged;

    // Staged per-integral operands: the batch is cut into chunks that run through a 3-stream
    // pipeline (H2D of chunk i+1 | kernel of chunk i | D2H of chunk i-1): PCIe is full duplex and
    // the integrals are independent.
    const int64_t kMinChunk = 1 << 16;
    static const int kMaxChunks = [] { const char* e = getenv("FTS_B200_CHUNKS"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 32 ? 32 : v); }();
    int nch = 1;
    if ((staged_in || staged_out) && (dim == 1 || batch >= 8 * kMinChunk)) nch = (int)(batch / kMinChunk < kMaxChunks ? batch / kMinChunk : kMaxChunks);
    if (nch < 1) nch = 1;
    // broadcast operands go first on the main stream; the pipeline streams wait for them
    for (Operand* o : ins)
        if (o->staged && o->elem == 0)
            FTS_CUDA(cudaMemcpyAsync(o->dev, o->host, (o == &op_ ? np : nb) * es, cudaMemcpyHostToDevice, g_stream));
    if (nch > 1) FTS_CUDA(cudaEventRecord(g_pipe_ev, g_stream));
    for (int ch = 0; ch < nch; ++ch) {
        cudaStream_t s = nch == 1 ? g_stream : g_pipe[ch % 3];
        const int64_t b0 = batch * ch / nch, b1 = batch * (ch + 1) / nch, nbatch = b1 - b0;
        if (nbatch == 0) continue;
        if (nch > 1) FTS_CUDA(cudaStreamWaitEvent(s, g_pipe_ev, 0));
        for (Operand* o : ins)
            if (o->staged && o->elem)
                FTS_CUDA(cudaMemcpyAsync(o->dev + (size_t)b0 * o->elem, (const char*)o->host + (size_t)b0 * o->elem, (size_t)nbatch * o->elem,
                                         cudaMemcpyHostToDevice, s));
        auto at = [&](const Operand& o) -> void* { return o.dev ? o.dev + (size_t)b0 * o.elem : nullptr; };
        if (int rc = fts_adaptive_batch_dev(c, f, order, options, at(ol), at(ou), bounds_stride, at(op_), params_stride, rtol, atol, max_levels, nbatch,
                                            at(ov), at(oe), (int32_t*)at(olv), (int32_t*)at(ofl), s))
            return rc;
        for (Operand* o : outs)
            if (o->staged)
                FTS_CUDA(cudaMemcpyAsync((char*)o->host + (size_t)b0 * o->elem, o->dev + (size_t)b0 * o->elem, (size_t)nbatch * o->elem,
                                         cudaMemcpyDeviceToHost, s));
    }
    tr_mark(1, tr_last);  // staging copies queued, kernels launched
    if (nch > 1)
        for (int k = 0; k < 3; ++k) FTS_CUDA(cudaStreamSynchronize(g_pipe[k]));
    FTS_CUDA(cudaStreamSynchronize(g_stream));
    tr_mark(2, tr_last);  // waiting for the device
    if (trace && ++tr_n == 1000) {
        fprintf(stderr, "[fts trace] fts_adaptive_batch, mean of 1000 calls: resolve %.2f us, queue %.2f us, wait %.2f us\n", tr_acc[0] / 1000, tr_acc[1] / 1000,
                tr_acc[2] / 1000);
        tr_n = 0;
        tr_acc[0] = tr_acc[1] = tr_acc[2] = 0;
    }
    return FTS_OK;
}

// ---- sharded single integral (multi-GPU: one process per GPU, exchange by the caller) -----------
template <class T>
static int sharded_level(fts_cache c, fts_integrand f, const void* low, const void* up, const void* params, int level, int rank, int nranks,
                         void* partial_dev, const void* state_dev, cudaStream_t s)
{
    const int dim = kind_dim(f->kind);
    const bool cmpl = kind_cmpl(f->kind);
    const void* fn = integrand_kernel(f, c->dtype, KID_LEVEL_GRID, 1);
    if (!fn) return set_error(FTS_ERR_UNSUPPORTED, "no whole-grid level kernel for this integrand/dtype");
    const size_t nblk = (size_t)g_prop.multiProcessorCount * 4;
    T* scratch = nullptr;
    FTS_CUDA(cudaMallocAsync((void**)&scratch, (2 * nblk + 2) * sizeof(T), s));
    unsigned int* ticket = (unsigned int*)(scratch + 2 * nblk);
    FTS_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), s));
    int rc = launch_level<T>(fn, c, dim, cmpl, (const T*)low, (const T*)up, (const T*)params, level, rank, nranks, scratch, ticket,
                             (T*)partial_dev, level ? (const T*)state_dev : nullptr, s);
    FTS_CUDA(cudaFreeAsync(scratch, s));
    return rc;
}

extern "C" int fts_adaptive_sharded_level_dev(fts_cache c, fts_integrand f, const void* low, const void* up, const void* params, int level,
                                              int rank, int nranks, void* partial_dev, const void* state_dev, void* stream)
{
    if (!c || !f || !low || !up || !partial_dev) return set_error(FTS_ERR_INVALID_ARGUMENT, "sharded level: null argument");
    if (int rc = check_cache_match(c, f)) return rc;
    if (kind_dim(f->kind) < 2) return set_error(FTS_ERR_UNSUPPORTED, "outer-axis sharding exists for 2D/3D integrals");
    if (level < 0 || level > c->max_levels) return set_error(FTS_ERR_INVALID_ARGUMENT, "provided adaptive cache supports " + std::to_string(c->max_levels) + " levels, but `max_levels=" + std::to_string(level) + "` was requested.");
    if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad rank/nranks");
    if (f->nparams > 0 && !params) return set_error(FTS_ERR_INVALID_ARGUMENT, "integrand needs parameters");
    if (int rc = ensure_init()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (c->dtype == FTS_F32) return sharded_level<float>(c, f, low, up, params, level, rank, nranks, partial_dev, state_dev, s);
    if (c->dtype == FTS_F64) return sharded_level<double>(c, f, low, up, params, level, rank, nranks, partial_dev, state_dev, s);
    return set_error(FTS_ERR_UNSUPPORTED, "sharded levels: f32/f64 only");
}

// ---- fused path: peer-memory mailboxes + ONE call that queues every level ------------------------
extern "C" int fts_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out)
{
    if (!dev_ptr || !ipc_handle_out || bytes == 0) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_peer_alloc: bad arguments");
    if (int rc = ensure_init()) return rc;
    FTS_CUDA(cudaMalloc(dev_ptr, bytes));
    FTS_CUDA(cudaMemset(*dev_ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    FTS_CUDA(cudaIpcGetMemHandle(&h, *dev_ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out, &h, sizeof h);
    return FTS_OK;
}
extern "C" int fts_peer_open(const void* ipc_handle, void** dev_ptr)
{
    if (!dev_ptr || !ipc_handle) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_peer_open: bad arguments");
    if (int rc = ensure_init()) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    FTS_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FTS_OK;
}
extern "C" int fts_peer_close(void* dev_ptr)
{
    if (dev_ptr) FTS_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return FTS_OK;
}
extern "C" int fts_peer_free(void* dev_ptr)
{
    if (dev_ptr) FTS_CUDA(cudaFree(dev_ptr));
    return FTS_OK;
}
// per-level exchange wait of the last fused call (ns, written by k_sharded_step_fused), and the
// spin timeout of that kernel
static unsigned long long peer_timeout_ns()
{
    static const unsigned long long v = [] {
        const char* e = getenv("FTS_B200_PEER_TIMEOUT_MS");
        const long long ms = e ? atoll(e) : 10000;
        return (unsigned long long)(ms < 1 ? 1 : ms) * 1000000ULL;
    }();
    return v;
}

extern "C" int fts_peer_mailbox_bytes(int nranks, size_t* bytes)
{
    if (nranks < 1 || nranks > FTS_MAX_RANKS || !bytes) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad nranks");
    *bytes = (size_t)(2 * 64 + 1) * nranks * 4 * sizeof(double);  // [epoch parity][level < 64][rank]{s, c, flag, pad}, then the row of fts_peer_barrier_dev
    return FTS_OK;
}

// Device-side barrier of the ranks' streams through the mailboxes: lane r raises this rank's counter in rank r's
// barrier row, then waits until every rank's counter in the own row has reached `count` (counts grow by one per
// barrier of the group).  Aligns the streams to a few microseconds where a host barrier leaves tens.
__global__ void k_peer_barrier(fts::PeerPtrs mail, int rank, int nranks, double count, unsigned long long timeout_ns, int* failed)
{
    const int lane = threadIdx.x;
    if (lane < nranks) {
        volatile double* theirs = mail.p[lane] + ((size_t)128 * nranks + rank) * 4;
        theirs[2] = count;
        __threadfence_system();
        const volatile double* mine = mail.p[rank] + ((size_t)128 * nranks + lane) * 4;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (mine[2] < count) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                if (failed) *failed = 1;
                break;
            }
        }
    }
}

extern "C" int fts_peer_barrier_dev(void* const* mailboxes, int rank, int nranks, uint64_t count, void* stream)
{
    if (!mailboxes || nranks < 1 || nranks > FTS_MAX_RANKS || rank < 0 || rank >= nranks || count == 0)
        return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_peer_barrier_dev: mailboxes, 0 <= rank < nranks <= FTS_MAX_RANKS and a non-zero count");
    if (int rc = ensure_init()) return rc;
    fts::PeerPtrs m;
    for (int r = 0; r < FTS_MAX_RANKS; ++r) m.p[r] = r < nranks ? (double*)mailboxes[r] : nullptr;
    k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(m, rank, nranks, (double)count, peer_timeout_ns(), nullptr);
    g_launches.fetch_add(1);
    FTS_CUDA(cudaGetLastError());
    return FTS_OK;
}

extern "C" int fts_peer_wait_ns(int nlevels, uint64_t* out)
{
    if (nlevels < 0 || nlevels > 64 || !out) return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_peer_wait_ns: 0 <= nlevels <= 64 and a buffer");
    if (!g_peer_wait) { memset(out, 0, sizeof(uint64_t) * nlevels); return FTS_OK; }
    FTS_CUDA(cudaMemcpy(out, g_peer_wait, sizeof(uint64_t) * nlevels, cudaMemcpyDeviceToHost));
    return FTS_OK;
}

template <class T>
static int sharded_fused(fts_cache c, fts_integrand f, const T* low, const T* up, const T* params, T rtol, T atol, int max_levels,
                         int shard_from, int rank, int nranks, void* const* mailboxes, double epoch, T* state, int options, cudaStream_t s)
{
    const int dim = kind_dim(f->kind);
    const bool cmpl = kind_cmpl(f->kind);
    const void* fn = integrand_kernel(f, c->dtype, KID_LEVEL_GRID, 1);
    if (!fn) return set_error(FTS_ERR_UNSUPPORTED, "no whole-grid level kernel for this integrand/dtype");
    const size_t nblk = (size_t)g_prop.multiProcessorCount * 4;
    T* scratch = nullptr;
    FTS_CUDA(cudaMallocAsync((void**)&scratch, (2 * nblk + 2 + 2) * sizeof(T) + 256, s));
    T* partial = scratch + 2 * nblk;
    unsigned int* ticket = (unsigned int*)(partial + 2);  // ticket, and 128 B behind it (its own line: polled by every CTA) the generation barrier
    unsigned int* gen = ticket + 32;
    FTS_CUDA(cudaMemsetAsync(ticket, 0, 33 * sizeof(unsigned int), s));
    if (!g_peer_wait) {
        FTS_CUDA(cudaMalloc((void**)&g_peer_wait, 64 * sizeof(unsigned long long)));
        FTS_CUDA(cudaMemsetAsync(g_peer_wait, 0, 64 * sizeof(unsigned long long), s));
    }
    const unsigned long long tmo = peer_timeout_ns();
    const T tm = load_elem<T>(c->tm);
    const int parity = ((long long)epoch) & 1;
    const bool one_launch = step_in_level_kernel();
    for (int level = 0; level <= max_levels; ++level) {
        const bool shard = mailboxes != nullptr && level >= shard_from;  // nranks == 1 with a mailbox exercises the flag protocol on one GPU
        fts::Args<T> a;
        memset(&a, 0, sizeof a);
        fill_cache_args(a, c);
        a.low = low; a.up = up; a.params = params; a.bstride = dim; a.pstride = 0; a.batch = 1;
        a.level = level; a.rank = shard ? rank : 0; a.nranks = shard ? nranks : 1;
        a.block_partials = scratch; a.ticket = ticket; a.partial_out = partial; a.state = level ? state : nullptr;
        if (shard) {
            for (int r = 0; r < nranks; ++r) a.peer_mail[r] = (double*)mailboxes[r];
            a.mail_slot = parity * 64 + level;
            a.epoch = epoch;
        }
        const unsigned blocks = grid_level_blocks(dim, level, a.nranks);
        if (one_launch) {
            // exchange wait + stop test in the last CTA of the level kernel; levels 0 ... range_top_level() in ONE launch
            int hi = level;
            if (level == 0) hi = max_levels < range_top_level() ? max_levels : range_top_level();
            a.rank = rank; a.nranks = mailboxes ? nranks : 1;
            a.shard_from = mailboxes ? shard_from : 0;
            if (mailboxes) {
                for (int r = 0; r < nranks; ++r) a.peer_mail[r] = (double*)mailboxes[r];
                a.mail_slot = parity * 64 + level;
                a.epoch = epoch;
            }
            a.state_out = state;
            a.rtol = rtol; a.atol = atol; a.options = options; a.max_levels = max_levels;
            a.timeout_ns = tmo; a.wait_ns = g_peer_wait;
            if (int rc = launch_level_range<T>(fn, a, dim, cmpl, level, hi, gen, s)) return rc;
            level = hi;
            continue;
        }
        if (int rc = launch_raw(fn, blocks, fts::CTA_THREADS, grid_level_smem(dim, cmpl, level, sizeof(T)), &a, s)) return rc;
        if (shard) {
            if (dim == 3) fts::k_sharded_step_fused<T, 3><<<1, 32, 0, s>>>(low, up, tm, rtol, atol, level, nranks, (const double*)mailboxes[rank], a.mail_slot, epoch, state, options, tmo, g_peer_wait, max_levels);
            else fts::k_sharded_step_fused<T, 2><<<1, 32, 0, s>>>(low, up, tm, rtol, atol, level, nranks, (const double*)mailboxes[rank], a.mail_slot, epoch, state, options, tmo, g_peer_wait, max_levels);
            g_launches.fetch_add(1);
        } else {
            if (dim == 3) launch_step<T, 3>(low, up, tm, rtol, atol, level, 1, partial, state, s, options, max_levels);
            else launch_step<T, 2>(low, up, tm, rtol, atol, level, 1, partial, state, s, options, max_levels);
        }
    }
    FTS_CUDA(cudaGetLastError());
    FTS_CUDA(cudaFreeAsync(scratch, s));
    return FTS_OK;
}

extern "C" int fts_adaptive_sharded_fused_dev(fts_cache c, fts_integrand f, const void* low, const void* up, const void* params,
                                              const void* rtol, const void* atol, int max_levels, int shard_from_level, int rank, int nranks,
                                              void* const* mailboxes, uint64_t epoch, void* state_dev, void* stream)
{
    return fts_adaptive_sharded_fused_opts_dev(c, f, low, up, params, rtol, atol, max_levels, shard_from_level, rank, nranks, mailboxes, epoch, state_dev, 0, stream);
}

extern "C" int fts_adaptive_sharded_fused_opts_dev(fts_cache c, fts_integrand f, const void* low, const void* up, const void* params,
                                                   const void* rtol, const void* atol, int max_levels, int shard_from_level, int rank, int nranks,
                                                   void* const* mailboxes, uint64_t epoch, void* state_dev, int options, void* stream)
{
    if (!c || !f || !low || !up || !rtol || !atol || !state_dev) return set_error(FTS_ERR_INVALID_ARGUMENT, "sharded fused: null argument");
    if (int rc = check_cache_match(c, f)) return rc;
    if (kind_dim(f->kind) < 2) return set_error(FTS_ERR_UNSUPPORTED, "outer-axis sharding exists for 2D/3D integrals");
    if (max_levels < 0 || max_levels > c->max_levels || max_levels >= 64)
        return set_error(FTS_ERR_INVALID_ARGUMENT, "provided adaptive cache supports " + std::to_string(c->max_levels) + " levels, but `max_levels=" + std::to_string(max_levels) + "` was requested.");
    if (nranks < 1 || nranks > FTS_MAX_RANKS || rank < 0 || rank >= nranks) return set_error(FTS_ERR_INVALID_ARGUMENT, "bad rank/nranks");
    if (nranks > 1 && (!mailboxes || epoch == 0)) return set_error(FTS_ERR_INVALID_ARGUMENT, "sharded fused: mailboxes and a non-zero epoch are required");
    if (f->nparams > 0 && !params) return set_error(FTS_ERR_INVALID_ARGUMENT, "integrand needs parameters");
    if (int rc = ensure_init()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (c->dtype == FTS_F64)
        return sharded_fused<double>(c, f, (const double*)low, (const double*)up, (const double*)params, load_elem<double>(rtol), load_elem<double>(atol),
                                     max_levels, shard_from_level, rank, nranks, mailboxes, (double)epoch, (double*)state_dev, options, s);
    if (c->dtype == FTS_F32)
        return sharded_fused<float>(c, f, (const float*)low, (const float*)up, (const float*)params, load_elem<float>(rtol), load_elem<float>(atol),
                                    max_levels, shard_from_level, rank, nranks, mailboxes, (double)epoch, (float*)state_dev, options, s);
    return set_error(FTS_ERR_UNSUPPORTED, "sharded fused: f32/f64 only");
}

extern "C" int fts_adaptive_sharded_step_dev(fts_cache c, const void* low, const void* up, const void* rtol, const void* atol, int level,
                                             int nranks, const void* partials_dev, void* state_dev, void* stream)
{
    return fts_adaptive_sharded_step_opts_dev(c, low, up, rtol, atol, level, nranks, partials_dev, state_dev, 0, 0, stream);
}

extern "C" int fts_adaptive_sharded_step_opts_dev(fts_cache c, const void* low, const void* up, const void* rtol, const void* atol, int level,
                                                  int nranks, const void* partials_dev, void* state_dev, int options, int max_levels, void* stream)
{
    if (!c || !low || !up || !rtol || !atol || !partials_dev || !state_dev) return set_error(FTS_ERR_INVALID_ARGUMENT, "sharded step: null argument");
    if (c->D != 2 && c->D != 3) return set_error(FTS_ERR_UNSUPPORTED, "sharded step needs a tensor cache");
    if (int rc = ensure_init()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (c->dtype == FTS_F64) {
        if (c->D == 3) launch_step<double, 3>((const double*)low, (const double*)up, load_elem<double>(c->tm), load_elem<double>(rtol), load_elem<double>(atol), level, nranks, (const double*)partials_dev, (double*)state_dev, s, options, max_levels);
        else launch_step<double, 2>((const double*)low, (const double*)up, load_elem<double>(c->tm), load_elem<double>(rtol), load_elem<double>(atol), level, nranks, (const double*)partials_dev, (double*)state_dev, s, options, max_levels);
    } else if (c->dtype == FTS_F32) {
        if (c->D == 3) launch_step<float, 3>((const float*)low, (const float*)up, load_elem<float>(c->tm), load_elem<float>(rtol), load_elem<float>(atol), level, nranks, (const float*)partials_dev, (float*)state_dev, s, options, max_levels);
        else launch_step<float, 2>((const float*)low, (const float*)up, load_elem<float>(c->tm), load_elem<float>(rtol), load_elem<float>(atol), level, nranks, (const float*)partials_dev, (float*)state_dev, s, options, max_levels);
    } else {
        return set_error(FTS_ERR_UNSUPPORTED, "sharded step: f32/f64 only");
    }
    FTS_CUDA(cudaGetLastError());
    return FTS_OK;
}

// ------------------------------------------------------------------------------- FMA peak ----
// Register-resident chains: 8 independent accumulators per thread, 64 FMAs per loop trip.
template <class T> __global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T a, T b)
{
    T r0 = a, r1 = a + T(1), r2 = a + T(2), r3 = a + T(3), r4 = a + T(4), r5 = a + T(5), r6 = a + T(6), r7 = a + T(7);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            r0 = fts::fm::fma(r0, b, a); r1 = fts::fm::fma(r1, b, a); r2 = fts::fm::fma(r2, b, a); r3 = fts::fm::fma(r3, b, a);
            r4 = fts::fm::fma(r4, b, a); r5 = fts::fm::fma(r5, b, a); r6 = fts::fm::fma(r6, b, a); r7 = fts::fm::fma(r7, b, a);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

extern "C" int fts_measure_fma_peak(int dtype, int iters, double* tflops_out, double* ms_out)
{
    if (int rc = ensure_init()) return rc;
    if (dtype != FTS_F32 && dtype != FTS_F64) return set_error(FTS_ERR_INVALID_ARGUMENT, "fma peak: f32 or f64");
    if (iters <= 0) iters = 4096;
    const int blocks = g_prop.multiProcessorCount * 8, threads = 256;
    void* d = nullptr;
    FTS_CUDA(cudaMalloc(&d, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    FTS_CUDA(cudaEventCreate(&e0));
    FTS_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        FTS_CUDA(cudaEventRecord(e0, g_stream));
        if (dtype == FTS_F64) k_fma_peak<double><<<blocks, threads, 0, g_stream>>>((double*)d, iters, 1.0000001, 0.9999999);
        else k_fma_peak<float><<<blocks, threads, 0, g_stream>>>((float*)d, iters, 1.0001f, 0.9999f);
        g_launches.fetch_add(1);
        FTS_CUDA(cudaEventRecord(e1, g_stream));
        FTS_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        FTS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    FTS_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    if (tflops_out) *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return FTS_OK;
}
