// fts_nvrtc.cpp — user integrands: CUDA-C++ body -> NVRTC -> sm_100a cubin -> same kernel templates.
//
// The reference specialises and inlines an arbitrary Julia callable into its loops
// (`f::F ... where F`, /root/reference/src/adaptive.jl:28).  A host closure cannot run on the
// GPU, so the boundary takes the integrand as source: the body is wrapped into a functor with
// the builtin-family protocol and the library's own kernel templates (fts_kernels.cuh, embedded
// in this library at build time) are instantiated for it.  Compilation needs no GPU; the cubin is
// loaded lazily on first launch.  Modules are cached by (body, kind, nparams).
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "fts_internal.h"

using namespace fts_host;

extern const char fts_src_dd[];       // fts_dd.cuh
extern const char fts_src_math[];     // fts_math.cuh      (generated: build/fts_embedded_src.cpp)
extern const char fts_src_kernels[];  // fts_kernels.cuh
extern const char fts_src_cta[];      // fts_kernels_cta.cuh
extern const char fts_src_exp_table[];  // fts_exp_table.inc
extern const char fts_src_log_table[];  // fts_log_table.inc
extern const char fts_src_dd_tables[];  // fts_dd_tables.inc

namespace {

struct KernelSpec {
    int dtype, kid, fast;
    std::string expr;
};

struct Program {  // one NVRTC compilation unit
    std::vector<char> cubin;
    std::vector<KernelSpec> specs;
    std::vector<std::string> lowered;
    cudaLibrary_t lib = nullptr;
    bool compiled = false, loaded = false;
};

struct Module {
    std::string body;
    int kind = 0, nparams = 0;
    Program base;  // float + double kernels, compiled when the integrand is created
    Program dd;    // Double64 kernels, compiled on first use (bodies may use functions without a dd overload)
    std::map<std::tuple<int, int, int>, const void*> kernels;
    std::mutex mu;
    int refs = 0;
};

std::mutex g_cache_mu;
std::map<std::tuple<std::string, int, int>, Module*> g_cache;

const char* kind_args(int kind)
{
    switch (kind) {
    case FTS_KIND_1D: return "T x";
    case FTS_KIND_2D: return "T x, T y";
    case FTS_KIND_3D: return "T x, T y, T z";
    case FTS_KIND_1D_CMPL: return "T x, T bmx, T xma";
    case FTS_KIND_2D_CMPL: return "T x, T y, T bmx, T xma, T dmy, T ymc";
    }
    if (kind > FTS_KIND_ND && kind <= FTS_KIND_ND + FTS_ND_MAX) return "const T* x";  // x[0] ... x[ND - 1]
    return nullptr;
}

void add_specs(std::vector<KernelSpec>& v, int kind, bool want_dd)
{
    const char* tn[3] = {"float", "double", "fts::dd"};
    const char *fixed = nullptr, *adapt = nullptr, *adapt_cta = nullptr, *fixed_cta = nullptr;
    switch (kind) {
    case FTS_KIND_1D: fixed = "k_fixed1d_tpi"; adapt = "k_adaptive1d_tpi"; adapt_cta = "k_adaptive1d_cta"; fixed_cta = "k_fixed1d_cta"; break;
    case FTS_KIND_2D: fixed = "k_fixed2d_tpi"; adapt = "k_adaptive2d_tpi"; adapt_cta = "k_adaptive2d_cta"; fixed_cta = "k_fixed2d_cta"; break;
    case FTS_KIND_3D: fixed = "k_fixed3d_tpi"; adapt = "k_adaptive3d_tpi"; adapt_cta = "k_adaptive3d_cta"; fixed_cta = "k_fixed3d_cta"; break;
    case FTS_KIND_1D_CMPL: adapt = "k_adaptive1d_cmpl_tpi"; adapt_cta = "k_adaptive1d_cmpl_cta"; break;
    case FTS_KIND_2D_CMPL: adapt = "k_adaptive2d_tpi"; adapt_cta = "k_adaptive2d_cta"; break;
    }
    if (kind > FTS_KIND_ND) {  // generic n-D tensor product: fixed grids, float / double
        if (want_dd) return;
        const std::string d = std::to_string(kind - FTS_KIND_ND);
        v.push_back({FTS_F32, KID_FIXED_ND, 1, "fts::k_fixed_nd<float, FtsUserIntegrand, " + d + ">"});
        v.push_back({FTS_F64, KID_FIXED_ND, 1, "fts::k_fixed_nd<double, FtsUserIntegrand, " + d + ">"});
        return;
    }
    const int ndt = 3;  // float, double, Double64 (2D/3D Double64: thread per integral only)
    for (int dt = 0; dt < ndt; ++dt) {
        if ((dt == FTS_DD64) != want_dd) continue;
        for (int fast = 0; fast < 2; ++fast) {
            if (dt == FTS_DD64 && fast) continue;
            const std::string targs = std::string("<") + tn[dt] + ", FtsUserIntegrand, " + (fast ? "true" : "false") + ">";
            if (fixed) v.push_back({dt, KID_FIXED_TPI, fast, std::string("fts::") + fixed + targs});
            if (adapt) v.push_back({dt, KID_ADAPT_TPI, fast, std::string("fts::") + adapt + targs});
            if (dt == FTS_DD64 && adapt && (kind == FTS_KIND_1D || kind == FTS_KIND_1D_CMPL))
                for (int g2 = 0; g2 < 2; ++g2)
                    v.push_back({dt, KID_ADAPT_GRP, g2, std::string("fts::k_adaptive1d_grp<fts::dd, FtsUserIntegrand, ") + (kind == FTS_KIND_1D_CMPL ? "true, " : "false, ") + (g2 ? "2>" : "4>")});
            if (kind == FTS_KIND_1D_CMPL && dt != FTS_DD64) v.push_back({dt, KID_FIXED_CMPL, fast, std::string("fts::k_fixed1d_cmpl_tpi") + targs});
            if (kind == FTS_KIND_1D || kind == FTS_KIND_1D_CMPL)
                v.push_back({dt, KID_ADAPT_TAIL, fast, std::string("fts::k_adaptive1d_tail<") + tn[dt] + ", FtsUserIntegrand, " + (fast ? "true" : "false") +
                                                           (kind == FTS_KIND_1D_CMPL ? ", true>" : ", false>")});
        }
        if (dt != FTS_DD64) {
            const std::string targs = std::string("<") + tn[dt] + ", FtsUserIntegrand>";
            if (adapt_cta) v.push_back({dt, KID_ADAPT_CTA, 1, std::string("fts::") + adapt_cta + targs});
            if (fixed_cta) v.push_back({dt, KID_FIXED_CTA, 1, std::string("fts::") + fixed_cta + targs});
            if (kind == FTS_KIND_2D || kind == FTS_KIND_2D_CMPL)
                v.push_back({dt, KID_LEVEL_GRID, 1, std::string("fts::k_level_grid<") + tn[dt] + ", FtsUserIntegrand, 2>"});
            if (kind == FTS_KIND_3D)
                v.push_back({dt, KID_LEVEL_GRID, 1, std::string("fts::k_level_grid<") + tn[dt] + ", FtsUserIntegrand, 3>"});
            if (kind == FTS_KIND_2D || kind == FTS_KIND_2D_CMPL) v.push_back({dt, KID_ADAPT_ORD, 0, std::string("fts::k_adaptive2d_ord") + targs});
            if (kind == FTS_KIND_2D || kind == FTS_KIND_2D_CMPL) v.push_back({dt, KID_ADAPT_WARP, 1, std::string("fts::k_adaptive2d_warp") + targs});
            if (kind == FTS_KIND_3D) v.push_back({dt, KID_ADAPT_ORD, 0, std::string("fts::k_adaptive3d_ord") + targs});
            if (kind == FTS_KIND_1D) v.push_back({dt, KID_FIXED_ORD, 0, std::string("fts::k_fixed1d_ord") + targs});
            if (kind == FTS_KIND_1D) v.push_back({dt, KID_EVAL, 1, std::string("fts::k_eval1d") + targs});
            if (kind == FTS_KIND_2D) v.push_back({dt, KID_FIXED_ORD, 0, std::string("fts::k_fixed2d_ord") + targs});
            if (kind == FTS_KIND_3D) v.push_back({dt, KID_FIXED_ORD, 0, std::string("fts::k_fixed3d_ord") + targs});
        }
    }
}

int nvrtc_fail(nvrtcResult r, const char* what, nvrtcProgram prog, char* log, size_t loglen)
{
    std::string msg = std::string("NVRTC: ") + nvrtcGetErrorString(r) + " in " + what;
    if (prog && log && loglen) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::vector<char> buf(n + 1, 0);
        if (n) nvrtcGetProgramLog(prog, buf.data());
        size_t m = n < loglen - 1 ? n : loglen - 1;
        memcpy(log, buf.data(), m);
        log[m] = 0;
    }
    return set_error(FTS_ERR_NVRTC, msg);
}

int compile(const std::string& body, int kind, int nparams, Program* m, bool want_dd, char* log, size_t loglen)
{
    const char* args = kind_args(kind);
    std::string src;
    src += "#include \"fts_kernels.cuh\"\n#include \"fts_kernels_cta.cuh\"\n";
    src += "struct FtsUserIntegrand {\n    static constexpr int NP = " + std::to_string(nparams) + ";\n";
    if (kind > FTS_KIND_ND) src += "    static constexpr int ND = " + std::to_string(kind - FTS_KIND_ND) + ";\n";
    src += std::string("    template <class T, bool FAST> static __device__ __forceinline__ T eval(") + args + ", const T* p)\n    {\n";
    src += body;
    src += "\n    }\n};\n";
    if (kind == FTS_KIND_2D_CMPL) src += "namespace fts { template <> struct is_cmpl2d<FtsUserIntegrand> { static constexpr bool value = true; }; }\n";

    const char* hdr_src[] = {fts_src_dd, fts_src_math, fts_src_kernels, fts_src_cta, fts_src_exp_table, fts_src_log_table, fts_src_dd_tables};
    const char* hdr_name[] = {"fts_dd.cuh", "fts_math.cuh", "fts_kernels.cuh", "fts_kernels_cta.cuh", "fts_exp_table.inc", "fts_log_table.inc", "fts_dd_tables.inc"};
    nvrtcProgram prog = nullptr;
    nvrtcResult r = nvrtcCreateProgram(&prog, src.c_str(), "fts_user_integrand.cu", 7, hdr_src, hdr_name);
    if (r != NVRTC_SUCCESS) return nvrtc_fail(r, "nvrtcCreateProgram", nullptr, log, loglen);
    add_specs(m->specs, kind, want_dd);
    for (auto& s : m->specs) {
        r = nvrtcAddNameExpression(prog, s.expr.c_str());
        if (r != NVRTC_SUCCESS) { nvrtc_fail(r, "nvrtcAddNameExpression", prog, log, loglen); nvrtcDestroyProgram(&prog); return FTS_ERR_NVRTC; }
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", "-default-device"};
    r = nvrtcCompileProgram(prog, 5, opts);
    if (r != NVRTC_SUCCESS) { nvrtc_fail(r, "nvrtcCompileProgram", prog, log, loglen); nvrtcDestroyProgram(&prog); return FTS_ERR_NVRTC; }
    for (auto& s : m->specs) {
        const char* low = nullptr;
        r = nvrtcGetLoweredName(prog, s.expr.c_str(), &low);
        if (r != NVRTC_SUCCESS) { nvrtc_fail(r, "nvrtcGetLoweredName", prog, log, loglen); nvrtcDestroyProgram(&prog); return FTS_ERR_NVRTC; }
        m->lowered.push_back(low);
    }
    size_t n = 0;
    r = nvrtcGetCUBINSize(prog, &n);
    if (r != NVRTC_SUCCESS || n == 0) { nvrtc_fail(r, "nvrtcGetCUBINSize", prog, log, loglen); nvrtcDestroyProgram(&prog); return FTS_ERR_NVRTC; }
    m->cubin.resize(n);
    r = nvrtcGetCUBIN(prog, m->cubin.data());
    nvrtcDestroyProgram(&prog);
    if (r != NVRTC_SUCCESS) return nvrtc_fail(r, "nvrtcGetCUBIN", nullptr, log, loglen);
    if (log && loglen) log[0] = 0;
    m->compiled = true;
    return FTS_OK;
}

bool load_program(Module* m, Program* pr)
{
    if (pr->loaded) return true;
    cudaError_t e = cudaLibraryLoadData(&pr->lib, pr->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { cuda_error(e, "cudaLibraryLoadData(user integrand cubin)"); return false; }
    for (size_t i = 0; i < pr->specs.size(); ++i) {
        cudaKernel_t k = nullptr;
        e = cudaLibraryGetKernel(&k, pr->lib, pr->lowered[i].c_str());
        if (e != cudaSuccess) { cuda_error(e, "cudaLibraryGetKernel"); return false; }
        m->kernels[std::make_tuple(pr->specs[i].dtype, pr->specs[i].kid, pr->specs[i].fast)] = (const void*)k;
    }
    pr->loaded = true;
    return true;
}

}  // namespace

// used by fts_host.cu
const void* fts_nvrtc_kernel(void* module, int dtype, int kid, int fast)
{
    Module* m = (Module*)module;
    std::lock_guard<std::mutex> lk(m->mu);
    Program* pr = dtype == FTS_DD64 ? &m->dd : &m->base;
    if (!pr->compiled) {
        char log[4096];
        if (compile(m->body, m->kind, m->nparams, pr, dtype == FTS_DD64, log, sizeof log)) {
            set_error(FTS_ERR_NVRTC, std::string("Double64 instantiation of the user integrand failed to compile:\n") + log);
            return nullptr;
        }
    }
    if (!load_program(m, pr)) return nullptr;
    auto it = m->kernels.find(std::make_tuple(dtype, kid, fast));
    if (it == m->kernels.end()) set_error(FTS_ERR_UNSUPPORTED, "no kernel of this style for the user integrand / dtype");
    return it == m->kernels.end() ? nullptr : it->second;
}

void fts_nvrtc_release(void* module)
{
    // modules stay in the process-wide cache (compiled once per distinct source)
    (void)module;
}

extern "C" int fts_integrand_nvrtc(const char* body, int kind, int nparams, fts_integrand* out, char* log, size_t loglen)
{
    if (!body || !out || !kind_args(kind) || nparams < 0 || nparams > 64)
        return set_error(FTS_ERR_INVALID_ARGUMENT, "fts_integrand_nvrtc: bad arguments");
    Module* m = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto key = std::make_tuple(std::string(body), kind, nparams);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) {
            m = it->second;
        } else {
            m = new Module();
            m->body = body; m->kind = kind; m->nparams = nparams;
            int rc = compile(body, kind, nparams, &m->base, false, log, loglen);
            if (rc) { delete m; return rc; }
            g_cache[key] = m;
        }
        m->refs++;
    }
    fts_integrand_s* f = new fts_integrand_s();
    f->builtin = 0; f->family = -1; f->kind = kind; f->nparams = nparams;
    f->module = m;
    *out = f;
    return FTS_OK;
}

// size of the cubin of an NVRTC integrand (lets CPU-only tests check that compilation happened)
extern "C" int fts_integrand_cubin_size(fts_integrand f, size_t* bytes, int* nkernels)
{
    if (!f || f->builtin || !f->module) return set_error(FTS_ERR_INVALID_ARGUMENT, "not an NVRTC integrand");
    Module* m = (Module*)f->module;
    if (bytes) *bytes = m->base.cubin.size();
    if (nkernels) *nkernels = (int)m->base.specs.size();
    return FTS_OK;
}
