// dtwa_api.cu -- the C ABI of libdicketwa.so (include/dicketwa.h) and its host-side logic:
// argument validation (the reference's error(...) calls), the fixed-step table, the reducer
// program, launch shapes, device arenas, the resampling rounds of tolerate_errors and the
// multi-device merge.  No CPU compute path exists here: without a CUDA device dtwa_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <thread>
#include <vector>

#include "dtwa_kernels.cuh"
#include "dtwa_expr.hpp"
#include "dtwa_jit.hpp"

using namespace dtwa;

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int set_err(int code, const std::string &m)
{
    g_err = m;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return set_err(DTWA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define TRY(call)            \
    do {                     \
        int rc_ = (call);    \
        if (rc_) return rc_; \
    } while (0)

extern "C" const char *dtwa_last_error(void) { return g_err.c_str(); }
extern "C" int dtwa_version(void) { return DTWA_VERSION; }

// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(DTWA_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct DevCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 0, clock_khz = 0, cc = 0;
    char name[256] = {0};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;                 // uploads overlapped with the ensemble kernel (host-supplied initial conditions)
    cudaEvent_t ev_copy_go = nullptr, ev_copy_done = nullptr;
    DevBuf partial, meta, u0, fail[2], scratch_a, scratch_b, scratch_c, ring;
    DevBuf arena_f64, arena_u64;   // result arenas of the ensemble in flight (grow-only: no cudaMalloc/cudaFree per call)
    // packed events left on the device by dtwa_integrate_events_packed (inside `ring`)
    int64_t packed_total = 0;
    int packed_nv = 0;
    double *packed_t = nullptr, *packed_u = nullptr;
    int32_t *packed_dir = nullptr;
};

// NCCL, loaded lazily (only a multi-device context needs it)
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
    void *handle = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

struct dtwa_ctx {
    std::vector<DevCtx> devs;
    int block_threads = 0, ctas_per_sm = 0;
    int jit_mode = 2;  // 0: interpret the reducer program, 1: always specialise with NVRTC, 2: specialise large ensembles
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    bool plan_live = false;   // a dtwa_plan of this context has not been freed yet: its arenas live in the DevCtx buffers
    // packed events of a split dtwa_integrate_events_packed: (device, first event, events) per block
    struct PackedBlock {
        int dev;
        int64_t first, count;
    };
    std::vector<PackedBlock> packed_blocks;
};

// ---- batch entry points on a multi-device context -------------------------------------------------------------------
// dtwa_integrate / _fm / _lyapunov / _integrate_events(_packed) / _shell_quadrature work on independent items with no
// exchange step.  On a context with several devices they give every device one contiguous block of the items (host
// arrays are addressed by block, nothing is copied or permuted) and run the single-device implementation for all
// blocks at once, one host thread per device.  g_dev selects the device of the calling thread; g_no_split marks the
// worker threads (and contexts with one device never split).
static thread_local int g_dev = 0;
static thread_local bool g_no_split = false;
static DevCtx &cur_dev(dtwa_ctx *ctx) { return ctx->devs[(size_t)g_dev < ctx->devs.size() ? g_dev : 0]; }
static bool should_split(const dtwa_ctx *ctx, int64_t n) { return ctx && !g_no_split && ctx->devs.size() > 1 && n >= 2 * (int64_t)ctx->devs.size(); }
static void store_max_ms(double *kernel_ms, const std::vector<double> &ms)
{
    if (kernel_ms) *kernel_ms = *std::max_element(ms.begin(), ms.end());
}
// call(k, lo, hi) -> DTWA_* for block k = items [lo, hi)
template <typename F>
static int split_over_devices(dtwa_ctx *ctx, int64_t n, F &&call)
{
    const int nd = (int)ctx->devs.size();
    const int64_t per = (n + nd - 1) / nd;
    std::vector<int> rc(nd, DTWA_OK);
    std::vector<std::string> msg(nd);
    std::vector<std::thread> th;
    for (int k = 0; k < nd; ++k) {
        const int64_t lo = std::min<int64_t>(n, k * per), hi = std::min<int64_t>(n, lo + per);
        if (hi <= lo) continue;
        th.emplace_back([&, k, lo, hi] {
            g_dev = k;
            g_no_split = true;
            rc[k] = call(k, lo, hi);
            if (rc[k] != DTWA_OK) msg[k] = g_err;
        });
    }
    for (auto &t : th) t.join();
    for (int k = 0; k < nd; ++k)
        if (rc[k] != DTWA_OK) return set_err(rc[k], msg[k]);
    return DTWA_OK;
}

static int load_nccl(dtwa_ctx *ctx)
{
    NcclApi &n = ctx->nccl;
    if (n.handle) return 0;
    const char *cands[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *c : cands) {
        n.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) break;
    }
    if (!n.handle) return set_err(DTWA_ERR_NCCL, std::string("cannot dlopen libnccl: ") + dlerror());
#define SYM(field, name)                                                                       \
    *(void **)(&n.field) = dlsym(n.handle, name);                                              \
    if (!n.field) return set_err(DTWA_ERR_NCCL, std::string("libnccl lacks symbol ") + name);
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return 0;
}

extern "C" int dtwa_create(dtwa_ctx **out, const int *device_ids, int ndev, void *stream)
{
    if (!out) return set_err(DTWA_ERR_INVALID, "dtwa_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_err(DTWA_ERR_NO_DEVICE,
                       std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0") +
                           "); libdicketwa has no CPU path");
    }
    if (ndev <= 0) ndev = 1;
    if (stream && ndev != 1) return set_err(DTWA_ERR_INVALID, "an external stream requires ndev == 1");
    dtwa_ctx *ctx = new dtwa_ctx;
    ctx->devs.resize(ndev);
    for (int i = 0; i < ndev; ++i) {
        DevCtx &d = ctx->devs[i];
        d.device = device_ids ? device_ids[i] : i;
        if (d.device < 0 || d.device >= count) {
            delete ctx;
            return set_err(DTWA_ERR_INVALID, "dtwa_create: device id out of range");
        }
        cudaDeviceProp prop;
        if (cudaSetDevice(d.device) != cudaSuccess || cudaGetDeviceProperties(&prop, d.device) != cudaSuccess) {
            delete ctx;
            return set_err(DTWA_ERR_CUDA, "cannot select device");
        }
        if (prop.major < 10) {
            delete ctx;
            return set_err(DTWA_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not sm_100-class; this library ships sm_100a code only");
        }
        d.sm_count = prop.multiProcessorCount;
        d.cc = prop.major * 10 + prop.minor;
        cudaDeviceGetAttribute(&d.clock_khz, cudaDevAttrClockRate, d.device);
        std::snprintf(d.name, sizeof(d.name), "%s", prop.name);
        if (stream) {
            d.stream = (cudaStream_t)stream;
            d.own_stream = false;
        } else if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return set_err(DTWA_ERR_CUDA, "cudaStreamCreate failed");
        }
        for (auto &ev : d.ev) cudaEventCreate(&ev);
        // (non-blocking: the caller's stream may be the legacy default stream, which a blocking stream would serialise with)
        if (cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking) != cudaSuccess) d.copy_stream = nullptr;
        cudaEventCreateWithFlags(&d.ev_copy_go, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&d.ev_copy_done, cudaEventDisableTiming);
    }
    if (ndev > 1) {
        int rc = load_nccl(ctx);
        if (rc) {
            delete ctx;
            return rc;
        }
        std::vector<int> ids(ndev);
        for (int i = 0; i < ndev; ++i) ids[i] = ctx->devs[i].device;
        ctx->comms.resize(ndev);
        int r = ctx->nccl.CommInitAll(ctx->comms.data(), ndev, ids.data());
        if (r) {
            std::string m = std::string("ncclCommInitAll: ") + ctx->nccl.GetErrorString(r);
            delete ctx;
            return set_err(DTWA_ERR_NCCL, m);
        }
    }
    *out = ctx;
    return DTWA_OK;
}

extern "C" int dtwa_destroy(dtwa_ctx *ctx)
{
    if (!ctx) return DTWA_OK;
    for (auto &c : ctx->comms)
        if (c) ctx->nccl.CommDestroy(c);
    for (auto &d : ctx->devs) {
        cudaSetDevice(d.device);
        cudaStreamSynchronize(d.stream);
        if (d.copy_stream) {
            cudaStreamSynchronize(d.copy_stream);
            cudaStreamDestroy(d.copy_stream);
        }
        if (d.ev_copy_go) cudaEventDestroy(d.ev_copy_go);
        if (d.ev_copy_done) cudaEventDestroy(d.ev_copy_done);
        for (DevBuf *b : {&d.partial, &d.meta, &d.u0, &d.fail[0], &d.fail[1], &d.scratch_a, &d.scratch_b, &d.scratch_c, &d.ring, &d.arena_f64, &d.arena_u64})
            b->release();
        for (auto &ev : d.ev)
            if (ev) cudaEventDestroy(ev);
        if (d.own_stream && d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
    return DTWA_OK;
}

extern "C" int dtwa_device_info(dtwa_ctx *ctx, int which, char *name, int *sm_count, int *clock_khz, int *cc)
{
    if (!ctx || which < 0 || which >= (int)ctx->devs.size()) return set_err(DTWA_ERR_INVALID, "bad context/device index");
    const DevCtx &d = ctx->devs[which];
    if (name) std::snprintf(name, 256, "%s", d.name);
    if (sm_count) *sm_count = d.sm_count;
    if (clock_khz) *clock_khz = d.clock_khz;
    if (cc) *cc = d.cc;
    return DTWA_OK;
}

extern "C" int dtwa_set_jit(dtwa_ctx *ctx, int mode)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (mode < 0 || mode > 2) return set_err(DTWA_ERR_INVALID, "jit mode must be 0 (interpreter), 1 (always) or 2 (auto)");
    ctx->jit_mode = mode;
    return DTWA_OK;
}

extern "C" int dtwa_jit_quiesce(void)
{
    jit_cache().quiesce();
    return DTWA_OK;
}

extern "C" int dtwa_set_launch(dtwa_ctx *ctx, int block_threads, int ctas_per_sm)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (block_threads < 0 || (block_threads % 32) != 0 || block_threads > 512)
        return set_err(DTWA_ERR_INVALID, "block_threads must be a multiple of 32 in [32, 512] (0 = default)");
    ctx->block_threads = block_threads;
    ctx->ctas_per_sm = ctas_per_sm < 0 ? 0 : ctas_per_sm;
    return DTWA_OK;
}

// ------------------------------------------------------------------------------------------------
// validation helpers
static int nvar_of(int kind) { return kind == DTWA_SYS_DICKE ? 4 : 2; }

static int check_system(const dtwa_system *sys)
{
    if (!sys) return set_err(DTWA_ERR_INVALID, "system is NULL");
    if (sys->kind != DTWA_SYS_DICKE && sys->kind != DTWA_SYS_LMG) return set_err(DTWA_ERR_INVALID, "unknown system kind");
    return 0;
}

static std::vector<std::string> varnames_of(int kind)
{
    if (kind == DTWA_SYS_DICKE) return {"Q", "q", "P", "p"};  // src/ClassicalDicke.jl:21
    return {"Q", "P"};                                         // src/ClassicalLMG.jl:15
}

static SysPar make_par(const dtwa_system *sys, int backwards)
{
    const double sg = backwards ? -1.0 : 1.0;
    SysPar p;
    if (sys->kind == DTWA_SYS_DICKE) {
        p.a = sg * sys->params[0];
        p.b = sg * sys->params[1];
        p.c = sg * sys->params[2];
        p.d = 2.0 * p.c;   // (the Bloch chart's 2 g)
    } else {
        const double Om = sg * sys->params[0], xi = sg * sys->params[1];
        p.a = Om;
        p.b = xi;
        p.c = xi / 2.0;
        p.d = 2.0 * xi + Om;
    }
    return p;
}

// which of the (sign-folded) Dicke parameters are exactly +-1, for the run-time specialisation (x * 1.0 == x exactly, so the
// specialised kernel keeps the bits of the generic one): (w == +-1 ? +-1 : 0) + 3 * (g == +-1 ? +-1 : 0), see jit_source
static int unit_flags(const dtwa_system *sys, int backwards)
{
    if (sys->kind != DTWA_SYS_DICKE) return 0;
    const SysPar p = make_par(sys, backwards);
    const int wf = p.b == 1.0 ? 1 : p.b == -1.0 ? -1 : 0;
    const int gf = p.c == 1.0 ? 1 : p.c == -1.0 ? -1 : 0;
    return wf + 3 * gf;
}

#ifndef DTWA_ADAPT_PAIR_DEFAULT
#define DTWA_ADAPT_PAIR_DEFAULT false
#endif
// Adaptive pairs with outputs by stepping onto ts[k]: the specialised kernel carries two warp streams per physical warp
// (adaptive_stream_pair).  DTWA_ADAPT_PAIR=0 in the environment selects the one-stream form (same results, bit for bit).
static bool pair_streams(const dtwa_solver *sol)
{
    static const bool enabled = [] {
        const char *e = std::getenv("DTWA_ADAPT_PAIR");
        return e ? std::atoi(e) != 0 : DTWA_ADAPT_PAIR_DEFAULT;
    }();
    const bool fixed = sol->alg == DTWA_ALG_RK4 || sol->alg == DTWA_ALG_RK8;
    return enabled && !fixed && sol->output == DTWA_OUTPUT_STEP;
}
// what jit_source specialises on: the unit flags, (+9) the two-stream form and (+18) the output mode of the adaptive kernels
static int jit_flags(const dtwa_system *sys, const dtwa_solver *sol)
{
    const bool fixed = sol->alg == DTWA_ALG_RK4 || sol->alg == DTWA_ALG_RK8;
    return unit_flags(sys, sol->backwards) + (pair_streams(sol) ? 9 : 0) + (!fixed && sol->output == DTWA_OUTPUT_HERMITE ? 18 : 0);
}

static int check_solver(const dtwa_solver *sol, bool fixed_needs_dt = true)
{
    if (!sol) return set_err(DTWA_ERR_INVALID, "solver is NULL");
    if (sol->alg < DTWA_ALG_RK4 || sol->alg > DTWA_ALG_TSIT5) return set_err(DTWA_ERR_INVALID, "unknown solver alg");
    const bool fixed = sol->alg == DTWA_ALG_RK4 || sol->alg == DTWA_ALG_RK8;
    if (fixed && fixed_needs_dt && !(sol->dt > 0)) return set_err(DTWA_ERR_INVALID, "fixed-step algorithms need dt > 0");
    if (!fixed && !(sol->tol > 0)) return set_err(DTWA_ERR_INVALID, "adaptive algorithms need tol > 0");
    if (sol->chart != DTWA_CHART_QP && sol->chart != DTWA_CHART_BLOCH) return set_err(DTWA_ERR_INVALID, "unknown chart");
    if (sol->output != DTWA_OUTPUT_STEP && sol->output != DTWA_OUTPUT_HERMITE) return set_err(DTWA_ERR_INVALID, "unknown output mode");
    return 0;
}

// Fixed-step table (same rule as oracle/oracle.c orc_step_table): interval k = (ts[k-1] or 0, ts[k]]
// in n = ceil(d/dt - 1e-9) equal sub-steps.  Also validates ts (src/ClassicalSystems.jl:89-91:
// "Use integate_backwards=true instead of negative time").
static int build_step_table(const double *ts, int nt, double dt, double wfold, std::vector<StepTab> &tab)
{
    if (!ts || nt < 1) return set_err(DTWA_ERR_INVALID, "ts must hold at least one time");
    tab.resize(nt);
    double tprev = 0.0;
    for (int k = 0; k < nt; ++k) {
        const double d = ts[k] - tprev;
        if (!(d >= 0)) {
            if (ts[k] < 0) return set_err(DTWA_ERR_INVALID, "Use integate_backwards=true instead of negative time");
            return set_err(DTWA_ERR_INVALID, "ts should be a sorted array of times");
        }
        StepTab r;
        std::memset(&r, 0, sizeof(r));
        if (d > 0 && dt > 0) {
            long long n = (long long)std::ceil(d / dt - 1e-9);
            if (n < 1) n = 1;
            if (n > 2000000000ll) return set_err(DTWA_ERR_INVALID, "dt too small for the requested ts spacing");
            r.n = (int32_t)n;
            r.h = d / (double)n;
            r.hh = r.h / 2.0;
            r.h3 = r.h / 3.0;
            r.h6 = r.h / 6.0;
            r.hw = r.h * wfold;
            r.hhw = r.hh * wfold;
            r.h3w = r.h3 * wfold;
            r.h6w = r.h6 * wfold;
        }
        tab[k] = r;
        tprev = ts[k];
    }
    return 0;
}

// Uniform grid detection (same rule as oracle/oracle.c orc_step_table): if every interval after the
// first has the same number of sub-steps and its length differs from the mean by no more than the
// rounding noise of the ts values themselves (8 ulp of the largest time), all those intervals use one
// sub-step h = (ts[nt-1] - ts[0]) / (n (nt-1)); interval 0 must be empty (ts[0] == 0) or of the same kind.
static void detect_uniform(const double *ts, int nt, const std::vector<StepTab> &tab, double wfold, SolverArgs &S)
{
    S.uniform = 0;
    S.first_n = 0;
    std::memset(&S.uni, 0, sizeof(S.uni));
    if (nt < 2 || tab[1].n < 1) return;
    const int n = tab[1].n;
    const double span = ts[nt - 1] - ts[0], mean = span / (double)(nt - 1);
    const double noise = 8.0 * 2.220446049250313e-16 * std::fabs(ts[nt - 1]);
    for (int k = 1; k < nt; ++k)
        if (tab[k].n != n || std::fabs((ts[k] - ts[k - 1]) - mean) > noise) return;
    if (!(tab[0].n == 0 || (tab[0].n == n && std::fabs(ts[0] - mean) <= noise))) return;
    StepTab r;
    std::memset(&r, 0, sizeof(r));
    r.n = n;
    r.h = mean / (double)n;
    r.hh = r.h / 2.0;
    r.h3 = r.h / 3.0;
    r.h6 = r.h / 6.0;
    r.hw = r.h * wfold;
    r.hhw = r.hh * wfold;
    r.h3w = r.h3 * wfold;
    r.h6w = r.h6 * wfold;
    S.uni = r;
    S.uniform = 1;
    S.first_n = tab[0].n;
}

// step-size controller of the adaptive pairs: OrdinaryDiffEq's defaults per algorithm (alg_utils.jl: beta2_default,
// beta1_default, qmin_default, qmax_default, gamma_default; qoldinit = 1e-4) unless the caller gives gains
static CtrlPar make_ctrl(const dtwa_solver *sol, bool force_scipy)
{
    CtrlPar c;
    std::memset(&c, 0, sizeof(c));
    const bool tsit5 = sol->alg == DTWA_ALG_TSIT5;       // generic gains of an order-5 method: beta2 = 2/(5 p), beta1 = 7/(10 p)
    const bool dp5 = sol->alg == DTWA_ALG_DP5 || tsit5;  // (what the two 5(4) pairs share: exponent 1/5, qmin 1/5, qmax 10)
    if (force_scipy || sol->controller == DTWA_CTRL_SCIPY) {
        c.kind = DTWA_CTRL_SCIPY;
        c.b1 = dp5 ? 0.2f : 0.125f;      // 1 / (error estimator order + 1)
        c.nb2 = 0.0f;
        c.g = 0.9f;
        c.fac_max = 10.0f;
        c.fac_min_acc = 0.0f;
        c.fac_min_rej = 0.2f;
        c.post_reject_cap = 1.0f;
        c.lqoldinit = 0.0f;
        c.acc_thr = 1.0;
        c.dom_shrink = 0.2;
        return c;
    }
    // gains: beta1 > 0 means the caller sets beta1 AND beta2 (beta2 = 0 is a legitimate choice); otherwise the defaults
    const double beta2 = sol->beta1 > 0 ? std::max(0.0, sol->beta2) : (tsit5 ? 0.08 : dp5 ? 0.04 : 0.0);
    const double beta1 = sol->beta1 > 0 ? sol->beta1 : (tsit5 ? 0.14 : dp5 ? 0.2 - 3.0 * beta2 / 4.0 : 0.125 - beta2 / 5.0);
    const double qmin = sol->qmin > 0 ? sol->qmin : (dp5 ? 0.2 : 1.0 / 3.0);
    const double qmax = sol->qmax > 0 ? sol->qmax : (dp5 ? 10.0 : 6.0);
    const double gamma = sol->gamma > 0 ? sol->gamma : 0.9;
    const double qoldinit = sol->qoldinit > 0 ? sol->qoldinit : 1e-4;
    c.kind = DTWA_CTRL_ODE;
    c.b1 = (float)beta1;
    c.nb2 = (float)(-beta2);
    c.g = (float)gamma;
    c.fac_max = (float)qmax;
    c.fac_min_acc = (float)qmin;
    c.fac_min_rej = (float)qmin;
    c.post_reject_cap = (float)qmax;
    c.lqoldinit = (float)std::log2(qoldinit);
    c.acc_thr = std::nextafter(1.0, 2.0);   // EEst <= 1
    c.dom_shrink = qmin;
    return c;
}

static SolverArgs make_solver_args(const dtwa_system *sys, const dtwa_solver *sol, const StepTab *d_tab, const double *d_ts,
                                   const double *ts, int nt, const std::vector<StepTab> &tab, bool force_scipy = false)
{
    SolverArgs S;
    S.ctrl = make_ctrl(sol, force_scipy);
    S.par = make_par(sys, sol->backwards);
    detect_uniform(ts, nt, tab, S.par.b, S);
    S.tab = d_tab;
    S.ts = d_ts;
    S.tol = sol->tol;
    S.dtmin = sol->dtmin > 0 ? sol->dtmin : sol->tol;  // src/ClassicalSystems.jl:138 dtmin=tol
    S.tend = ts[nt - 1];
    S.maxiters = sol->maxiters > 0 ? sol->maxiters : 100000;
    S.nt = nt;
    S.interp = (!force_scipy && sol->output == DTWA_OUTPUT_HERMITE && sol->alg != DTWA_ALG_RK4 && sol->alg != DTWA_ALG_RK8) ? 1 : 0;
    return S;
}

static int make_sampler_par(const dtwa_system *sys, const dtwa_sampler *smp, SamplerPar &sp)
{
    if (!smp) return set_err(DTWA_ERR_INVALID, "sampler is NULL");
    const int nv = nvar_of(sys->kind);
    sp.kind = smp->kind;
    sp.nv = nv;
    for (int i = 0; i < 4; ++i) sp.center[i] = smp->center[i];
    sp.hbar = smp->hbar;
    sp.seed = smp->seed;
    sp.offset = smp->offset;
    sp.u0 = nullptr;
    switch (smp->kind) {
    case DTWA_SAMPLER_HOST_ARRAY:
    case DTWA_SAMPLER_DEVICE_ARRAY:
        if (!smp->u0) return set_err(DTWA_ERR_INVALID, "array sampler needs u0");
        return 0;
    case DTWA_SAMPLER_COHERENT_HWXSU2:
        if (nv != 4) return set_err(DTWA_ERR_INVALID, "coherent_Wigner_HWxSU2 samples 4 variables; the system has 2");
        break;
    case DTWA_SAMPLER_COHERENT_HW:
    case DTWA_SAMPLER_COHERENT_SU2:
        if (nv != 2) return set_err(DTWA_ERR_INVALID, "coherent_Wigner_HW / _SU2 sample 2 variables; the system has 4");
        break;
    default: return set_err(DTWA_ERR_INVALID, "unknown sampler kind");
    }
    if (!(smp->hbar > 0)) return set_err(DTWA_ERR_INVALID, "sampler needs hbar > 0");
    return 0;
}

template <typename F>
static int by_system(int kind, F &&f)
{
    if (kind == DTWA_SYS_DICKE) return f(std::integral_constant<int, DTWA_SYS_DICKE>());
    if (kind == DTWA_SYS_DICKE_BLOCH) return f(std::integral_constant<int, DTWA_SYS_DICKE_BLOCH>());
    return f(std::integral_constant<int, DTWA_SYS_LMG>());
}
template <typename F>
static int by_base_system(int kind, F &&f)   // entry points that only exist in the reference's chart
{
    if (kind == DTWA_SYS_DICKE) return f(std::integral_constant<int, DTWA_SYS_DICKE>());
    return f(std::integral_constant<int, DTWA_SYS_LMG>());
}
// the template system the integrating kernels are instantiated for: the caller's system in the chart dtwa_solver.chart names
static int chart_sys(const dtwa_system *sys, const dtwa_solver *sol)
{
    return (sys->kind == DTWA_SYS_DICKE && sol->chart == DTWA_CHART_BLOCH) ? DTWA_SYS_DICKE_BLOCH : sys->kind;
}
template <typename F>
static int by_alg(int alg, F &&f)
{
    switch (alg) {
    case DTWA_ALG_RK4: return f(std::integral_constant<int, DTWA_ALG_RK4>());
    case DTWA_ALG_RK8: return f(std::integral_constant<int, DTWA_ALG_RK8>());
    case DTWA_ALG_DP5: return f(std::integral_constant<int, DTWA_ALG_DP5>());
    case DTWA_ALG_TSIT5: return f(std::integral_constant<int, DTWA_ALG_TSIT5>());
    default: return f(std::integral_constant<int, DTWA_ALG_DOP853>());
    }
}

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------
// simple batch entry points
extern "C" int dtwa_hamiltonian(dtwa_ctx *ctx, const dtwa_system *sys, const double *u, int64_t n, double *h)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    TRY(check_system(sys));
    if (n < 0 || (n > 0 && (!u || !h))) return set_err(DTWA_ERR_INVALID, "bad buffers");
    if (n == 0) return DTWA_OK;
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    const int nv = nvar_of(sys->kind);
    TRY(d.scratch_a.ensure(sizeof(double) * n * nv));
    TRY(d.scratch_b.ensure(sizeof(double) * n));
    CK(cudaMemcpyAsync(d.scratch_a.p, u, sizeof(double) * n * nv, cudaMemcpyHostToDevice, d.stream));
    if (sys->kind == DTWA_SYS_DICKE)
        hamiltonian_kernel<DTWA_SYS_DICKE><<<blocks_for(n, 256), 256, 0, d.stream>>>(sys->params[0], sys->params[1], sys->params[2],
                                                                                  (const double *)d.scratch_a.p, n, (double *)d.scratch_b.p);
    else
        hamiltonian_kernel<DTWA_SYS_LMG><<<blocks_for(n, 256), 256, 0, d.stream>>>(sys->params[0], sys->params[1], 0.0,
                                                                                (const double *)d.scratch_a.p, n, (double *)d.scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h, d.scratch_b.p, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

extern "C" int dtwa_sample(dtwa_ctx *ctx, const dtwa_sampler *smp, int64_t n, uint32_t attempt, double *u)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (!smp || n < 0 || (n > 0 && !u)) return set_err(DTWA_ERR_INVALID, "bad arguments");
    if (smp->kind < DTWA_SAMPLER_COHERENT_HW || smp->kind > DTWA_SAMPLER_COHERENT_HWXSU2)
        return set_err(DTWA_ERR_INVALID, "dtwa_sample needs a coherent_Wigner_* sampler");
    if (!(smp->hbar > 0)) return set_err(DTWA_ERR_INVALID, "sampler needs hbar > 0");
    if (n == 0) return DTWA_OK;
    dtwa_system fake;
    fake.kind = smp->kind == DTWA_SAMPLER_COHERENT_HWXSU2 ? DTWA_SYS_DICKE : DTWA_SYS_LMG;
    SamplerPar sp;
    TRY(make_sampler_par(&fake, smp, sp));
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    TRY(d.scratch_a.ensure(sizeof(double) * n * sp.nv));
    if (sp.nv == 4)
        sample_kernel<4><<<blocks_for(n, 128), 128, 0, d.stream>>>(sp, n, attempt, (double *)d.scratch_a.p);
    else
        sample_kernel<2><<<blocks_for(n, 128), 128, 0, d.stream>>>(sp, n, attempt, (double *)d.scratch_a.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(u, d.scratch_a.p, sizeof(double) * n * sp.nv, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

extern "C" int dtwa_pdf(dtwa_ctx *ctx, const dtwa_sampler *smp, const double *u, int64_t n, double *w)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (!smp || n < 0 || (n > 0 && (!u || !w))) return set_err(DTWA_ERR_INVALID, "bad arguments");
    if (smp->kind < DTWA_SAMPLER_COHERENT_HW || smp->kind > DTWA_SAMPLER_COHERENT_HWXSU2)
        return set_err(DTWA_ERR_INVALID, "dtwa_pdf needs a coherent_Wigner_* sampler");
    if (n == 0) return DTWA_OK;
    dtwa_system fake;
    fake.kind = smp->kind == DTWA_SAMPLER_COHERENT_HWXSU2 ? DTWA_SYS_DICKE : DTWA_SYS_LMG;
    SamplerPar sp;
    TRY(make_sampler_par(&fake, smp, sp));
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    TRY(d.scratch_a.ensure(sizeof(double) * n * sp.nv));
    TRY(d.scratch_b.ensure(sizeof(double) * n));
    CK(cudaMemcpyAsync(d.scratch_a.p, u, sizeof(double) * n * sp.nv, cudaMemcpyHostToDevice, d.stream));
    if (sp.nv == 4)
        pdf_kernel<4><<<blocks_for(n, 128), 128, 0, d.stream>>>(sp, (const double *)d.scratch_a.p, n, (double *)d.scratch_b.p);
    else
        pdf_kernel<2><<<blocks_for(n, 128), 128, 0, d.stream>>>(sp, (const double *)d.scratch_a.p, n, (double *)d.scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(w, d.scratch_b.p, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

extern "C" int dtwa_eval_expr(dtwa_ctx *ctx, const dtwa_system *sys, const char *expr, const double *u, int64_t n, double *v)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    TRY(check_system(sys));
    if (!expr || n < 0 || (n > 0 && (!u || !v))) return set_err(DTWA_ERR_INVALID, "bad arguments");
    Program prog;
    try {
        auto vars = varnames_of(sys->kind);
        ExprCompiler c(vars);
        c.compile(expr, prog);
    } catch (const std::exception &e) {
        return set_err(DTWA_ERR_EXPR, e.what());
    }
    if (prog.max_depth > VM_STACK) return set_err(DTWA_ERR_EXPR, "expression too deeply nested for the VM stack");
    prog.emit(OP_STORE);
    prog.emit(OP_END);
    if (n == 0) return DTWA_OK;
    if (prog.consts.empty()) prog.consts.push_back(0.0);
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    const int nv = nvar_of(sys->kind);
    const size_t pb = sizeof(VmIns) * prog.ins.size(), cb = sizeof(double) * prog.consts.size();
    TRY(d.scratch_a.ensure(sizeof(double) * n * nv));
    TRY(d.scratch_b.ensure(sizeof(double) * n));
    TRY(d.scratch_c.ensure(((cb + 15) & ~size_t(15)) + pb));
    unsigned char *meta = (unsigned char *)d.scratch_c.p;
    CK(cudaMemcpyAsync(meta, prog.consts.data(), cb, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(meta + ((cb + 15) & ~size_t(15)), prog.ins.data(), pb, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.scratch_a.p, u, sizeof(double) * n * nv, cudaMemcpyHostToDevice, d.stream));
    const VmIns *dprog = (const VmIns *)(meta + ((cb + 15) & ~size_t(15)));
    if (nv == 4)
        eval_kernel<4><<<blocks_for(n, 128), 128, 0, d.stream>>>(dprog, (const double *)meta, (const double *)d.scratch_a.p, n, (double *)d.scratch_b.p);
    else
        eval_kernel<2><<<blocks_for(n, 128), 128, 0, d.stream>>>(dprog, (const double *)meta, (const double *)d.scratch_a.p, n, (double *)d.scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(v, d.scratch_b.p, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

static int make_range(const dtwa_range &r, RangeD &out)
{
    if (r.len < 1) return set_err(DTWA_ERR_INVALID, "histogram range needs len >= 1");
    if (!(r.stop >= r.start)) return set_err(DTWA_ERR_INVALID, "histogram range needs stop >= start");
    out.start = r.start;
    out.stop = r.stop;
    out.len = r.len;
    out.lenm1 = (double)(r.len - 1);
    out.c = out.lenm1 / (r.stop - r.start);   // (Inf/NaN for a degenerate range: the fast path then never accepts)
    out.pad = 0.0;
    return 0;
}

extern "C" int dtwa_scale_to_index(dtwa_ctx *ctx, const double *v, int64_t n, const dtwa_range *r, int64_t *idx)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (!r || n < 0 || (n > 0 && (!v || !idx))) return set_err(DTWA_ERR_INVALID, "bad arguments");
    RangeD rd;
    TRY(make_range(*r, rd));
    if (n == 0) return DTWA_OK;
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    TRY(d.scratch_a.ensure(sizeof(double) * n));
    TRY(d.scratch_b.ensure(sizeof(long long) * n));
    CK(cudaMemcpyAsync(d.scratch_a.p, v, sizeof(double) * n, cudaMemcpyHostToDevice, d.stream));
    bin_kernel<<<blocks_for(n, 256), 256, 0, d.stream>>>((const double *)d.scratch_a.p, n, rd, (long long *)d.scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(idx, d.scratch_b.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

// ------------------------------------------------------------------------------------------------
// a7: batch integrate
extern "C" int dtwa_integrate(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n,
                              const double *ts, int32_t nt, double *out, int32_t *status, int64_t *nsteps)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    TRY(check_system(sys));
    TRY(check_solver(sol));
    if (n < 0 || (n > 0 && (!u0 || !out || !status))) return set_err(DTWA_ERR_INVALID, "bad buffers");
    std::vector<StepTab> tab;
    TRY(build_step_table(ts, nt, sol->dt, make_par(sys, sol->backwards).b, tab));
    if (n == 0) return DTWA_OK;
    const int nv = nvar_of(sys->kind);
    if (should_split(ctx, n))
        return split_over_devices(ctx, n, [&](int, int64_t lo, int64_t hi) {
            return dtwa_integrate(ctx, sys, sol, u0 + lo * nv, hi - lo, ts, nt, out + (size_t)lo * nt * nv, status + lo, nsteps ? nsteps + lo : nullptr);
        });
    DevCtx &d = cur_dev(ctx);
    CK(cudaSetDevice(d.device));
    const size_t tab_b = sizeof(StepTab) * nt, ts_b = sizeof(double) * nt;
    TRY(d.meta.ensure(tab_b + ts_b));
    TRY(d.u0.ensure(sizeof(double) * n * nv));
    TRY(d.scratch_a.ensure(sizeof(double) * (size_t)n * nt * nv));
    TRY(d.scratch_b.ensure(sizeof(int32_t) * n));
    TRY(d.scratch_c.ensure(sizeof(long long) * n));
    unsigned char *meta = (unsigned char *)d.meta.p;
    CK(cudaMemcpyAsync(meta, tab.data(), tab_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(meta + tab_b, ts, ts_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.u0.p, u0, sizeof(double) * n * nv, cudaMemcpyHostToDevice, d.stream));
    const SolverArgs S = make_solver_args(sys, sol, (const StepTab *)meta, (const double *)(meta + tab_b), ts, nt, tab);
    const unsigned grid = blocks_for(n, 128);
    if (sol->chart == DTWA_CHART_BLOCH && sys->kind != DTWA_SYS_DICKE)
        return set_err(DTWA_ERR_INVALID, "chart = DTWA_CHART_BLOCH is defined for the Dicke model (the LMG field is polynomial in (Q,P) already)");
    by_system(chart_sys(sys, sol), [&](auto sk) {
        return by_alg(sol->alg, [&](auto ak) {
            integrate_kernel<decltype(sk)::value, decltype(ak)::value><<<grid, 128, 0, d.stream>>>(
                S, (const double *)d.u0.p, n, (double *)d.scratch_a.p, (int32_t *)d.scratch_b.p, (long long *)d.scratch_c.p);
            return 0;
        });
    });
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d.scratch_a.p, sizeof(double) * (size_t)n * nt * nv, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(status, d.scratch_b.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d.stream));
    if (nsteps) CK(cudaMemcpyAsync(nsteps, d.scratch_c.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

// ------------------------------------------------------------------------------------------------
// fundamental matrix and Lyapunov spectra (src/ClassicalSystems.jl:110-128,183-224)
template <typename F>
static int by_var_system(int kind, F &&f)
{
    if (kind == DTWA_SYS_DICKE) return f(std::integral_constant<int, DTWA_SYS_DICKE_VAR>());
    return f(std::integral_constant<int, DTWA_SYS_LMG_VAR>());
}

extern "C" int dtwa_integrate_fm(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0,
                                 const double *fm0, int64_t n, const double *ts, int32_t nt, double *out_u, double *out_fm,
                                 int32_t *status, int64_t *nsteps)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    TRY(check_system(sys));
    TRY(check_solver(sol));
    if (n < 0 || (n > 0 && (!u0 || !out_u || !out_fm || !status))) return set_err(DTWA_ERR_INVALID, "bad buffers");
    std::vector<StepTab> tab;
    TRY(build_step_table(ts, nt, sol->dt, make_par(sys, sol->backwards).b, tab));
    if (n == 0) return DTWA_OK;
    const int nv = nvar_of(sys->kind), nn = nv * nv;
    if (should_split(ctx, n))
        return split_over_devices(ctx, n, [&](int, int64_t lo, int64_t hi) {
            return dtwa_integrate_fm(ctx, sys, sol, u0 + lo * nv, fm0 ? fm0 + lo * nn : nullptr, hi - lo, ts, nt, out_u + (size_t)lo * nt * nv,
                                     out_fm + (size_t)lo * nt * nn, status + lo, nsteps ? nsteps + lo : nullptr);
        });
    DevCtx &d = cur_dev(ctx);
    CK(cudaSetDevice(d.device));
    const size_t tab_b = sizeof(StepTab) * nt, ts_b = sizeof(double) * nt;
    const size_t b_u0 = sizeof(double) * n * nv, b_fm0 = fm0 ? sizeof(double) * n * nn : 0;
    const size_t b_ou = sizeof(double) * (size_t)n * nt * nv, b_of = sizeof(double) * (size_t)n * nt * nn;
    TRY(d.meta.ensure(tab_b + ts_b));
    TRY(d.u0.ensure(b_u0 + b_fm0));
    TRY(d.scratch_a.ensure(b_ou + b_of));
    TRY(d.scratch_b.ensure(sizeof(int32_t) * n));
    TRY(d.scratch_c.ensure(sizeof(long long) * n));
    unsigned char *meta = (unsigned char *)d.meta.p;
    CK(cudaMemcpyAsync(meta, tab.data(), tab_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(meta + tab_b, ts, ts_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.u0.p, u0, b_u0, cudaMemcpyHostToDevice, d.stream));
    if (fm0) CK(cudaMemcpyAsync((unsigned char *)d.u0.p + b_u0, fm0, b_fm0, cudaMemcpyHostToDevice, d.stream));
    const SolverArgs S = make_solver_args(sys, sol, (const StepTab *)meta, (const double *)(meta + tab_b), ts, nt, tab, /*force_scipy=*/true);
    const unsigned grid = blocks_for(n * nv, 128);
    const double *d_fm0 = fm0 ? (const double *)((unsigned char *)d.u0.p + b_u0) : nullptr;
    by_var_system(sys->kind, [&](auto sk) {
        return by_alg(sol->alg, [&](auto ak) {
            integrate_fm_kernel<decltype(sk)::value, decltype(ak)::value><<<grid, 128, 0, d.stream>>>(
                S, (const double *)d.u0.p, d_fm0, n, (double *)d.scratch_a.p, (double *)((unsigned char *)d.scratch_a.p + b_ou),
                (int32_t *)d.scratch_b.p, (long long *)d.scratch_c.p);
            return 0;
        });
    });
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_u, d.scratch_a.p, b_ou, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(out_fm, (unsigned char *)d.scratch_a.p + b_ou, b_of, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(status, d.scratch_b.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d.stream));
    if (nsteps) CK(cudaMemcpyAsync(nsteps, d.scratch_c.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

extern "C" int dtwa_lyapunov(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n,
                             double t_interval, double lam_tol, double lam_reltol, int32_t maxiters, double *lambda,
                             double *x_final, int32_t *iters, int32_t *status, int64_t *nsteps, double *kernel_ms)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    TRY(check_system(sys));
    TRY(check_solver(sol));
    if (n < 0 || (n > 0 && (!u0 || !lambda || !iters || !status))) return set_err(DTWA_ERR_INVALID, "bad buffers");
    if (!(t_interval > 0)) return set_err(DTWA_ERR_INVALID, "the renormalisation interval t must be > 0");
    if (maxiters < 1) return set_err(DTWA_ERR_INVALID, "maxiters must be >= 1");
    std::vector<StepTab> tab;
    const double ts1[1] = {t_interval};
    TRY(build_step_table(ts1, 1, sol->dt, make_par(sys, sol->backwards).b, tab));
    if (n == 0) return DTWA_OK;
    const int nv = nvar_of(sys->kind);
    if (should_split(ctx, n)) {
        // The cost of an orbit (rounds until convergence) varies smoothly over a Lyapunov map, so contiguous blocks of a
        // map are unbalanced.  Inputs and outputs are a few numbers per orbit: deal the orbits round-robin (orbit i to
        // device i mod ndev) through permuted host copies and scatter the results back.
        const int64_t nd = (int64_t)ctx->devs.size();
        std::vector<int64_t> src;   // permuted position -> original index: residue classes mod ndev, one after the other
        src.reserve((size_t)n);     // (the blocks of split_over_devices cut this list within a few items of the class ends)
        for (int64_t k = 0; k < nd; ++k)
            for (int64_t i = k; i < n; i += nd) src.push_back(i);
        std::vector<double> pu0((size_t)n * nv), plam((size_t)n * nv), px(x_final ? (size_t)n * nv : 0);
        std::vector<int32_t> pit((size_t)n), pst((size_t)n);
        std::vector<int64_t> pns(nsteps ? (size_t)n : 0);
        for (int64_t p = 0; p < n; ++p)
            for (int c = 0; c < nv; ++c) pu0[(size_t)p * nv + c] = u0[src[(size_t)p] * nv + c];
        std::vector<double> ms((size_t)nd, 0.0);
        TRY(split_over_devices(ctx, n, [&](int k, int64_t lo, int64_t hi) {
            return dtwa_lyapunov(ctx, sys, sol, pu0.data() + lo * nv, hi - lo, t_interval, lam_tol, lam_reltol, maxiters, plam.data() + lo * nv,
                                 x_final ? px.data() + lo * nv : nullptr, pit.data() + lo, pst.data() + lo, nsteps ? pns.data() + lo : nullptr,
                                 &ms[k]);
        }));
        for (int64_t p = 0; p < n; ++p) {
            const int64_t i = src[(size_t)p];
            for (int c = 0; c < nv; ++c) {
                lambda[i * nv + c] = plam[(size_t)p * nv + c];
                if (x_final) x_final[i * nv + c] = px[(size_t)p * nv + c];
            }
            iters[i] = pit[(size_t)p];
            status[i] = pst[(size_t)p];
            if (nsteps) nsteps[i] = pns[(size_t)p];
        }
        store_max_ms(kernel_ms, ms);
        return DTWA_OK;
    }
    DevCtx &d = cur_dev(ctx);
    CK(cudaSetDevice(d.device));
    const size_t tab_b = sizeof(StepTab), ts_b = sizeof(double);
    const size_t b_u0 = sizeof(double) * n * nv;
    TRY(d.meta.ensure(tab_b + ts_b));
    TRY(d.u0.ensure(b_u0));
    TRY(d.scratch_a.ensure(2 * b_u0));
    TRY(d.scratch_b.ensure(2 * sizeof(int32_t) * n));
    TRY(d.scratch_c.ensure(sizeof(long long) * (n + 1)));
    unsigned char *meta = (unsigned char *)d.meta.p;
    CK(cudaMemcpyAsync(meta, tab.data(), tab_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(meta + tab_b, ts1, ts_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.u0.p, u0, b_u0, cudaMemcpyHostToDevice, d.stream));
    const SolverArgs S = make_solver_args(sys, sol, (const StepTab *)meta, (const double *)(meta + tab_b), ts1, 1, tab, /*force_scipy=*/true);
    LyapArgs L;
    L.t_interval = t_interval;
    L.ltol = lam_tol;
    L.lreltol = lam_reltol;
    L.maxiters = maxiters;
    L.pad = 0;
    double *d_lam = (double *)d.scratch_a.p, *d_x = (double *)((unsigned char *)d.scratch_a.p + b_u0);
    int32_t *d_it = (int32_t *)d.scratch_b.p, *d_st = d_it + n;
    // persistent grid: groups draw initial conditions from a counter (kept behind the nsteps array)
    unsigned long long *d_next = (unsigned long long *)d.scratch_c.p + n;
    CK(cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), d.stream));
    CK(cudaEventRecord(d.ev[0], d.stream));
    by_var_system(sys->kind, [&](auto sk) {
        return by_alg(sol->alg, [&](auto ak) {
            auto kern = lyapunov_kernel<decltype(sk)::value, decltype(ak)::value>;
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
            const long long want = (n * nv + 127) / 128, cap = (long long)d.sm_count * per_sm;
            const unsigned grid = (unsigned)std::max(1ll, std::min(want, cap));
            kern<<<grid, 128, 0, d.stream>>>(S, L, (const double *)d.u0.p, n, d_lam, d_x, d_it, d_st, (long long *)d.scratch_c.p, d_next);
            return 0;
        });
    });
    CK(cudaGetLastError());
    CK(cudaEventRecord(d.ev[3], d.stream));
    CK(cudaMemcpyAsync(lambda, d_lam, b_u0, cudaMemcpyDeviceToHost, d.stream));
    if (x_final) CK(cudaMemcpyAsync(x_final, d_x, b_u0, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(iters, d_it, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(status, d_st, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d.stream));
    if (nsteps) CK(cudaMemcpyAsync(nsteps, d.scratch_c.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    if (kernel_ms) {
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, d.ev[0], d.ev[3]));
        *kernel_ms = m;
    }
    return DTWA_OK;
}

// ------------------------------------------------------------------------------------------------
// surfaces of section (integrate + ContinuousCallback, docs/src/ClassicalDickeExamples.md:82-109)
// Runs the section kernel and leaves the padded event arrays on the device (scratch_a: t | u, scratch_b: dir | count |
// status); counts, end states, status and step counts are copied to the host.  The stream is synchronised on return.
struct EventsOnDevice {
    double *d_t = nullptr, *d_u = nullptr;
    int32_t *d_dir = nullptr, *d_cnt = nullptr;
    int nv = 0;
    size_t cap = 0;
};
static int events_solve(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n, double t_end,
                        const dtwa_event *ev, int32_t max_events, int32_t *ev_count, double *u_end, int32_t *status, int64_t *nsteps,
                        double *kernel_ms, EventsOnDevice &out)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    TRY(check_system(sys));
    TRY(check_solver(sol));
    if (!ev) return set_err(DTWA_ERR_INVALID, "NULL event");
    if (n < 0 || max_events < 0 || (n > 0 && (!u0 || !ev_count || !status || !u_end))) return set_err(DTWA_ERR_INVALID, "bad buffers");
    if (!(t_end >= 0)) return set_err(DTWA_ERR_INVALID, "Use integate_backwards=true instead of negative time");
    if (ev->direction < -1 || ev->direction > 1) return set_err(DTWA_ERR_INVALID, "event direction must be -1, 0 or +1");
    const int nv = nvar_of(sys->kind);
    double cn = 0;
    for (int i = 0; i < nv; ++i) cn += std::fabs(ev->coef[i]);
    if (!(cn > 0) || !std::isfinite(cn) || !std::isfinite(ev->offset))
        return set_err(DTWA_ERR_INVALID, "the event function coef . u - offset needs a non-zero, finite coefficient vector");
    std::vector<StepTab> tab;
    const double ts1[1] = {t_end};
    TRY(build_step_table(ts1, 1, sol->dt, make_par(sys, sol->backwards).b, tab));
    out.nv = nv;
    out.cap = (size_t)max_events;
    if (n == 0) return DTWA_OK;
    DevCtx &d = cur_dev(ctx);
    CK(cudaSetDevice(d.device));
    const size_t tab_b = sizeof(StepTab), ts_b = sizeof(double);
    const size_t b_u0 = sizeof(double) * n * nv, cap = (size_t)max_events;
    const size_t b_t = sizeof(double) * n * cap, b_u = b_t * nv, b_dir = sizeof(int32_t) * n * cap;
    TRY(d.meta.ensure(tab_b + ts_b));
    TRY(d.u0.ensure(2 * b_u0));
    TRY(d.scratch_a.ensure(b_t + b_u + 16));
    TRY(d.scratch_b.ensure(b_dir + 2 * sizeof(int32_t) * n + 16));
    TRY(d.scratch_c.ensure(sizeof(long long) * n));
    unsigned char *meta = (unsigned char *)d.meta.p;
    CK(cudaMemcpyAsync(meta, tab.data(), tab_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(meta + tab_b, ts1, ts_b, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.u0.p, u0, b_u0, cudaMemcpyHostToDevice, d.stream));
    const SolverArgs S = make_solver_args(sys, sol, (const StepTab *)meta, (const double *)(meta + tab_b), ts1, 1, tab, /*force_scipy=*/true);
    EventArgs E;
    std::memset(&E, 0, sizeof(E));
    for (int i = 0; i < nv; ++i) E.coef[i] = ev->coef[i];
    E.offset = ev->offset;
    E.t_end = t_end;
    E.direction = ev->direction;
    E.cap = max_events;
    double *d_t = (double *)d.scratch_a.p, *d_u = d_t + n * cap;
    int32_t *d_dir = (int32_t *)d.scratch_b.p, *d_cnt = d_dir + n * cap, *d_st = d_cnt + n;
    double *d_end = (double *)((unsigned char *)d.u0.p + b_u0);
    CK(cudaEventRecord(d.ev[0], d.stream));
    by_base_system(sys->kind, [&](auto sk) {
        return by_alg(sol->alg, [&](auto ak) {
            events_kernel<decltype(sk)::value, decltype(ak)::value><<<blocks_for(n, EVENTS_BLOCK), EVENTS_BLOCK, 0, d.stream>>>(
                S, E, (const double *)d.u0.p, n, d_t, d_u, d_dir, d_cnt, d_end, d_st, (long long *)d.scratch_c.p);
            return 0;
        });
    });
    CK(cudaGetLastError());
    CK(cudaEventRecord(d.ev[3], d.stream));
    CK(cudaMemcpyAsync(ev_count, d_cnt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(status, d_st, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(u_end, d_end, b_u0, cudaMemcpyDeviceToHost, d.stream));
    if (nsteps) CK(cudaMemcpyAsync(nsteps, d.scratch_c.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    if (kernel_ms) {
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, d.ev[0], d.ev[3]));
        *kernel_ms = m;
    }
    out.d_t = d_t;
    out.d_u = d_u;
    out.d_dir = d_dir;
    out.d_cnt = d_cnt;
    return DTWA_OK;
}

extern "C" int dtwa_integrate_events(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n,
                                     double t_end, const dtwa_event *ev, int32_t max_events, double *ev_t, double *ev_u,
                                     int32_t *ev_dir, int32_t *ev_count, double *u_end, int32_t *status, int64_t *nsteps,
                                     double *kernel_ms)
{
    if (n > 0 && max_events > 0 && (!ev_t || !ev_u || !ev_dir)) return set_err(DTWA_ERR_INVALID, "bad buffers");
    if (should_split(ctx, n) && sys && max_events >= 0 && u0 && ev_count && status && u_end) {
        const size_t nv = (size_t)nvar_of(sys->kind), cap = (size_t)max_events;
        std::vector<double> ms(ctx->devs.size(), 0.0);
        TRY(split_over_devices(ctx, n, [&](int k, int64_t lo, int64_t hi) {
            return dtwa_integrate_events(ctx, sys, sol, u0 + lo * nv, hi - lo, t_end, ev, max_events, ev_t ? ev_t + lo * cap : nullptr,
                                         ev_u ? ev_u + lo * cap * nv : nullptr, ev_dir ? ev_dir + lo * cap : nullptr, ev_count + lo,
                                         u_end + lo * nv, status + lo, nsteps ? nsteps + lo : nullptr, &ms[k]);
        }));
        store_max_ms(kernel_ms, ms);
        return DTWA_OK;
    }
    EventsOnDevice E;
    TRY(events_solve(ctx, sys, sol, u0, n, t_end, ev, max_events, ev_count, u_end, status, nsteps, kernel_ms, E));
    if (n == 0 || E.cap == 0) return DTWA_OK;
    DevCtx &d = cur_dev(ctx);
    const size_t cap = E.cap, nv = (size_t)E.nv;
    // only the columns some trajectory used travel back (the arrays are [n][max_events] padded)
    size_t used = 0;
    for (int64_t i = 0; i < n; ++i) used = std::max(used, (size_t)std::min<int64_t>(ev_count[i], (int64_t)cap));
    if (used) {
        CK(cudaMemcpy2DAsync(ev_t, cap * sizeof(double), E.d_t, cap * sizeof(double), used * sizeof(double), n, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaMemcpy2DAsync(ev_u, cap * nv * sizeof(double), E.d_u, cap * nv * sizeof(double), used * nv * sizeof(double), n,
                             cudaMemcpyDeviceToHost, d.stream));
        CK(cudaMemcpy2DAsync(ev_dir, cap * sizeof(int32_t), E.d_dir, cap * sizeof(int32_t), used * sizeof(int32_t), n,
                             cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
    }
    return DTWA_OK;
}

// Packed (CSR) event output: row i owns the entries [offsets[i], offsets[i+1]) of ev_t / ev_u / ev_dir, offsets[i+1] -
// offsets[i] = min(ev_count[i], max_events).  Two calls, because the caller can only size its buffers once the total
// is known: _packed solves, compacts on the device and returns the total; _packed_fetch copies the compacted arrays.
extern "C" int dtwa_integrate_events_packed(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0,
                                            int64_t n, double t_end, const dtwa_event *ev, int32_t max_events, int64_t *offsets,
                                            int32_t *ev_count, double *u_end, int32_t *status, int64_t *nsteps, double *kernel_ms)
{
    if (!offsets) return set_err(DTWA_ERR_INVALID, "bad buffers");
    if (should_split(ctx, n) && sys && max_events >= 0 && u0 && ev_count && status && u_end) {
        // every device packs its block (offsets local to the block), then the offsets are made global
        const size_t nv = (size_t)nvar_of(sys->kind);
        const int nd = (int)ctx->devs.size();
        std::vector<double> ms(nd, 0.0);
        std::vector<int64_t> lo_of(nd, -1), hi_of(nd, -1);
        std::vector<std::vector<int64_t>> loc(nd);
        ctx->packed_blocks.clear();
        TRY(split_over_devices(ctx, n, [&](int k, int64_t lo, int64_t hi) {
            lo_of[k] = lo;
            hi_of[k] = hi;
            loc[k].assign((size_t)(hi - lo) + 1, 0);
            return dtwa_integrate_events_packed(ctx, sys, sol, u0 + lo * nv, hi - lo, t_end, ev, max_events, loc[k].data(), ev_count + lo,
                                                u_end + lo * nv, status + lo, nsteps ? nsteps + lo : nullptr, &ms[k]);
        }));
        int64_t first = 0;
        offsets[0] = 0;
        for (int k = 0; k < nd; ++k) {
            if (lo_of[k] < 0) continue;
            for (int64_t i = lo_of[k]; i < hi_of[k]; ++i) offsets[i + 1] = first + loc[k][(size_t)(i - lo_of[k]) + 1];
            const int64_t cnt = loc[k].back();
            ctx->packed_blocks.push_back({k, first, cnt});
            first += cnt;
        }
        store_max_ms(kernel_ms, ms);
        return DTWA_OK;
    }
    EventsOnDevice E;
    offsets[0] = 0;
    if (ctx) {
        cur_dev(ctx).packed_total = 0;
        if (!g_no_split) ctx->packed_blocks.clear();
    }
    TRY(events_solve(ctx, sys, sol, u0, n, t_end, ev, max_events, ev_count, u_end, status, nsteps, kernel_ms, E));
    if (n == 0) return DTWA_OK;
    for (int64_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + std::min<int64_t>(ev_count[i], (int64_t)E.cap);
    const int64_t total = offsets[n];
    DevCtx &d = cur_dev(ctx);
    d.packed_nv = E.nv;
    if (total == 0) return DTWA_OK;
    const size_t nv = (size_t)E.nv;
    const size_t b_off = sizeof(int64_t) * (size_t)(n + 1), b_t = sizeof(double) * (size_t)total;
    TRY(d.ring.ensure(b_off + b_t * (1 + nv) + sizeof(int32_t) * (size_t)total + 64));
    int64_t *d_off = (int64_t *)d.ring.p;
    double *p_t = (double *)((unsigned char *)d.ring.p + ((b_off + 15) & ~(size_t)15)), *p_u = p_t + total;
    int32_t *p_dir = (int32_t *)(p_u + (size_t)total * nv);
    CK(cudaMemcpyAsync(d_off, offsets, b_off, cudaMemcpyHostToDevice, d.stream));
    const long long warps_per_block = 8;
    pack_events_kernel<<<(unsigned)((n + warps_per_block - 1) / warps_per_block), 32 * warps_per_block, 0, d.stream>>>(
        E.d_t, E.d_u, E.d_dir, d_off, n, (long long)E.cap, E.nv, p_t, p_u, p_dir);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(d.stream));
    d.packed_total = total;
    d.packed_t = p_t;
    d.packed_u = p_u;
    d.packed_dir = p_dir;
    return DTWA_OK;
}

extern "C" int dtwa_integrate_events_packed_fetch(dtwa_ctx *ctx, int64_t total, double *ev_t, double *ev_u, int32_t *ev_dir)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (!g_no_split && !ctx->packed_blocks.empty()) {   // the result of a split solve: one segment per device
        int64_t sum = 0;
        for (const auto &b : ctx->packed_blocks) sum += b.count;
        if (total != sum) return set_err(DTWA_ERR_INVALID, "no packed events of this size are pending (call dtwa_integrate_events_packed first)");
        if (total == 0) return DTWA_OK;
        if (!ev_t || !ev_u || !ev_dir) return set_err(DTWA_ERR_INVALID, "bad buffers");
        const auto blocks = ctx->packed_blocks;
        std::vector<int> rc(blocks.size(), DTWA_OK);
        std::vector<std::string> msg(blocks.size());
        std::vector<std::thread> th;
        for (size_t i = 0; i < blocks.size(); ++i)
            th.emplace_back([&, i] {
                const auto &b = blocks[i];
                g_dev = b.dev;
                g_no_split = true;
                const size_t nv = (size_t)ctx->devs[b.dev].packed_nv;
                rc[i] = dtwa_integrate_events_packed_fetch(ctx, b.count, ev_t + b.first, ev_u + (size_t)b.first * nv, ev_dir + b.first);
                if (rc[i] != DTWA_OK) msg[i] = g_err;
            });
        for (auto &t : th) t.join();
        for (size_t i = 0; i < blocks.size(); ++i)
            if (rc[i] != DTWA_OK) return set_err(rc[i], msg[i]);
        return DTWA_OK;
    }
    DevCtx &d = cur_dev(ctx);
    if (total != d.packed_total) return set_err(DTWA_ERR_INVALID, "no packed events of this size are pending (call dtwa_integrate_events_packed first)");
    if (total == 0) return DTWA_OK;
    if (!ev_t || !ev_u || !ev_dir) return set_err(DTWA_ERR_INVALID, "bad buffers");
    CK(cudaSetDevice(d.device));
    CK(cudaMemcpyAsync(ev_t, d.packed_t, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(ev_u, d.packed_u, sizeof(double) * (size_t)total * d.packed_nv, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(ev_dir, d.packed_dir, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    return DTWA_OK;
}

// ------------------------------------------------------------------------------------------------
// energy-shell quadrature (src/EnergyShellProjections.jl:49-96) for classical expressions
extern "C" int dtwa_shell_quadrature(dtwa_ctx *ctx, const dtwa_system *sys, double eps, const double *QP, int64_t ncell,
                                     const char *const *exprs, int32_t nexpr, double p_res, int32_t onlyqroot, double *out,
                                     int32_t *valid, int32_t *nodes, double *kernel_ms)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    TRY(check_system(sys));
    if (sys->kind != DTWA_SYS_DICKE) return set_err(DTWA_ERR_INVALID, "the energy-shell quadrature is defined for the Dicke model");
    if (ncell < 0 || (ncell > 0 && (!QP || !out || !valid))) return set_err(DTWA_ERR_INVALID, "bad buffers");
    if (!exprs || nexpr < 1 || nexpr > SHELL_MAX_EXPR) return set_err(DTWA_ERR_INVALID, "between 1 and 8 expressions are required");
    if (!(p_res > 0)) return set_err(DTWA_ERR_INVALID, "p_res must be > 0");
    if (onlyqroot < -1 || onlyqroot > 1) return set_err(DTWA_ERR_INVALID, "onlyqroot must be -1, 0 or +1");
    Program prog;
    std::vector<int32_t> offs;
    try {
        auto vars = varnames_of(sys->kind);
        for (int e = 0; e < nexpr; ++e) {
            if (!exprs[e]) return set_err(DTWA_ERR_INVALID, "NULL expression");
            offs.push_back((int32_t)prog.ins.size());
            ExprCompiler c(vars);
            c.compile(exprs[e], prog);
            prog.emit(OP_STORE);
            prog.emit(OP_END);
        }
    } catch (const std::exception &e) {
        return set_err(DTWA_ERR_EXPR, e.what());
    }
    if (prog.max_depth > VM_STACK) return set_err(DTWA_ERR_EXPR, "expression too deeply nested for the VM stack");
    if (ncell == 0) return DTWA_OK;
    if (should_split(ctx, ncell)) {
        std::vector<double> ms(ctx->devs.size(), 0.0);
        TRY(split_over_devices(ctx, ncell, [&](int k, int64_t lo, int64_t hi) {
            return dtwa_shell_quadrature(ctx, sys, eps, QP + 2 * lo, hi - lo, exprs, nexpr, p_res, onlyqroot, out + lo * nexpr, valid + lo,
                                         nodes ? nodes + lo : nullptr, &ms[k]);
        }));
        store_max_ms(kernel_ms, ms);
        return DTWA_OK;
    }
    if (prog.consts.empty()) prog.consts.push_back(0.0);
    DevCtx &d = cur_dev(ctx);
    CK(cudaSetDevice(d.device));
    const size_t cb = (sizeof(double) * prog.consts.size() + 15) & ~size_t(15), pb = (sizeof(VmIns) * prog.ins.size() + 15) & ~size_t(15);
    const size_t ob = sizeof(int32_t) * offs.size();
    TRY(d.meta.ensure(cb + pb + ob));
    TRY(d.u0.ensure(sizeof(double) * 2 * ncell));
    TRY(d.scratch_a.ensure(sizeof(double) * ncell * nexpr));
    TRY(d.scratch_b.ensure(2 * sizeof(int32_t) * ncell));
    unsigned char *meta = (unsigned char *)d.meta.p;
    CK(cudaMemcpyAsync(meta, prog.consts.data(), sizeof(double) * prog.consts.size(), cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(meta + cb, prog.ins.data(), sizeof(VmIns) * prog.ins.size(), cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(meta + cb + pb, offs.data(), ob, cudaMemcpyHostToDevice, d.stream));
    CK(cudaMemcpyAsync(d.u0.p, QP, sizeof(double) * 2 * ncell, cudaMemcpyHostToDevice, d.stream));
    ShellArgs A;
    A.w0 = sys->params[0];
    A.w = sys->params[1];
    A.g = sys->params[2];
    A.eps = eps;
    A.p_res = p_res;
    A.onlyqroot = onlyqroot;
    A.nexpr = nexpr;
    A.consts = (const double *)meta;
    A.prog = (const VmIns *)(meta + cb);
    A.prog_off = (const int32_t *)(meta + cb + pb);
    int32_t *d_valid = (int32_t *)d.scratch_b.p, *d_nodes = d_valid + ncell;
    CK(cudaEventRecord(d.ev[0], d.stream));
    shell_quadrature_kernel<<<blocks_for(ncell * 32, 128), 128, 0, d.stream>>>(A, (const double *)d.u0.p, ncell, (double *)d.scratch_a.p,
                                                                              d_valid, d_nodes);
    CK(cudaGetLastError());
    CK(cudaEventRecord(d.ev[3], d.stream));
    CK(cudaMemcpyAsync(out, d.scratch_a.p, sizeof(double) * ncell * nexpr, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(valid, d_valid, sizeof(int32_t) * ncell, cudaMemcpyDeviceToHost, d.stream));
    if (nodes) CK(cudaMemcpyAsync(nodes, d_nodes, sizeof(int32_t) * ncell, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    if (kernel_ms) {
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, d.ev[0], d.ev[3]));
        *kernel_ms = m;
    }
    return DTWA_OK;
}

// ------------------------------------------------------------------------------------------------
// a13-a20: the ensemble
struct RedLayout {
    int kind = 0;
    int nexpr = 0;
    int slot0 = 0;          // first sums slot (AVERAGE / SURVIVAL)
    int64_t u64_off = 0;    // first counts element in the u64 arena (HIST)
    int64_t counts_len = 0;
};

struct DevPlan {
    DevCtx *dc = nullptr;
    long long lo = 0, n = 0;  // this device's trajectory block
    double *d_f64 = nullptr;
    unsigned long long *d_u64 = nullptr;
    int grid = 0, block = 0;
};

struct dtwa_plan {
    dtwa_ctx *ctx = nullptr;
    int nt = 0, nslots = 0;
    int64_t n_f64 = 0, n_u64 = 0;
    std::vector<RedLayout> reds;
    std::vector<DevPlan> devs;
    int launches = 0, rounds = 0;
    bool fetched = false, jit = false;
};

struct LaunchShape {
    int block, grid, window;   // grid: CTAs -- for the adaptive kernels the number of warp STREAMS (per_warp of them per CTA)
    int per_warp = 1;
    size_t smem;
    double *ring;   // adaptive kernels: ring of raw states in global memory, [grid][window][32 lanes][nvar]
};

// DTWA_ADAPT_WARPS=n (environment; tuning experiment): the run-time specialised adaptive kernels are compiled for and launched
// with n warp streams per CTA (dtwa_kernels.cuh); the interpreter kernels keep one.
static int adapt_warps_env()
{
    static const int v = [] {
        const char *e = std::getenv("DTWA_ADAPT_WARPS");
        const int n = e ? std::atoi(e) : 1;
        return (n == 2 || n == 4 || n == 8) ? n : 1;
    }();
    return v;
}

// Ring length of the adaptive kernels (output times per warp stream).  A long ring lets the lanes of a warp drift apart (lane
// efficiency 0.97 at 256 against 0.94 at 64); a short one costs less per output -- measured on a B200 (profiles/r2_ring_length.md):
// with about one output per attempted step (configs[0]: ts = 0:0.05:100 at tol 1e-12) 64 output times are 16 % faster for an
// average, 5 % for histograms and the same for a variance, with one output per 3 or more attempts 256 are 1-2 % faster.  The
// attempts per output interval are estimated from the tolerance: the embedded pairs take steps of about tol^(1/p) time units of
// the fastest frequency (p = 8: 0.032 at 1e-12, 0.1 at 1e-8, measured 0.033 / 0.09; p = 5: measured 1.2-1.4 x that).
static int adaptive_ring_cap(const dtwa_system *sys, const dtwa_solver *sol, const double *ts, int nt)
{
    if (nt < 2 || !(sol->tol > 0)) return 256;
    double wmax = 1.0;
    const int npar = sys->kind == DTWA_SYS_DICKE ? 3 : 2;
    for (int i = 0; i < npar; ++i) wmax = std::max(wmax, std::fabs(sys->params[i]));
    const double h = std::pow(sol->tol, (sol->alg == DTWA_ALG_DP5 || sol->alg == DTWA_ALG_TSIT5) ? 0.2 : 0.125) / wmax;
    const double spacing = std::fabs(ts[nt - 1] - ts[0]) / (nt - 1);
    return (h >= 0.5 * spacing) ? 64 : 256;
}

static int shape_for(dtwa_ctx *ctx, DevCtx &d, const void *kern, int alg, const EnsArgs &A, long long n, int nvar, LaunchShape &ls, int per_warp = 1,
                     int warps = 1, int ring_cap = 256)
{
    const bool fixed = alg == DTWA_ALG_RK4 || alg == DTWA_ALG_RK8;
    const int max_threads = alg == DTWA_ALG_RK4 ? ens_launch<DTWA_ALG_RK4>::max_threads : alg == DTWA_ALG_RK8 ? ens_launch<DTWA_ALG_RK8>::max_threads : 32;
    int block = ctx->block_threads ? ctx->block_threads : 256;
    if (block > max_threads) block = max_threads;
    if (!fixed) block = 32 * warps;   // adaptive kernels: independent warp streams (warp-level sliding flush, see ensemble_body), one per CTA by default
    // window: output times staged between two flushes.  Fixed step: one row per warp in shared memory (<= 24 KB),
    // one CTA-wide flush per window (16 -> 32 outputs per barrier is worth 0.6 %, measured).  Adaptive: a ring of raw
    // states in global memory (32 lanes x nvar doubles per output time and CTA), long enough (<= 256 output times, <= 1 GiB)
    // that lanes only wait for each other when they drift more than half a ring apart; shared memory holds one row of sums.
    const size_t slots = A.nslots > 0 ? A.nslots : 1;
    const int rows = fixed ? block / 32 : block;   // rows of the shared-memory staging (adaptive: one row of sums per lane for the ring flush)
    const size_t per_out = sizeof(double) * rows * slots + sizeof(unsigned) * A.stage_bins + sizeof(unsigned);
    size_t wcap = fixed ? 32 : (size_t)ring_cap, wbudget = 24 * 1024;
    if (const char *e = std::getenv("DTWA_WINDOW")) wcap = (size_t)std::max(1, std::atoi(e));  // tuning experiments
    int window = (int)std::max<size_t>(1, std::min<size_t>({wcap, (size_t)A.S.nt, fixed ? wbudget / per_out : wcap}));
    size_t smem = ens_smem_bytes(A.ncst, A.nrange, A.nhist, A.nslots, A.stage_bins, A.nins, rows, fixed ? window : 1, A.stage2d_bins);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem));
    if (per_sm < 1) return set_err(DTWA_ERR_CAPACITY, "ensemble kernel does not fit on an SM with this reducer set");
    if (ctx->ctas_per_sm > 0 && ctx->ctas_per_sm < per_sm) per_sm = ctx->ctas_per_sm;
    long long want = fixed ? (n + block - 1) / block : (n + 31) / 32;
    long long cap = (long long)d.sm_count * per_sm * per_warp * (fixed ? 1 : warps);
    const int grid = (int)std::max(1ll, std::min(want, cap));
    ls.per_warp = per_warp * (fixed ? 1 : warps);   // warp streams per CTA
    ls.ring = nullptr;
    if (!fixed) {
        // ring length: even, <= wcap, and the whole ring buffer <= 1 GiB
        const size_t per_ring_out = sizeof(double) * 32 * (size_t)nvar * (size_t)grid;
        window = (int)std::max<size_t>(1, std::min<size_t>((size_t)window, ((size_t)1 << 30) / per_ring_out));
        if (window > 1) window &= ~1;
        TRY(d.ring.ensure(per_ring_out * (size_t)window));
        d.packed_total = 0;   // (packed events of an earlier dtwa_integrate_events_packed lived here)
        ls.ring = (double *)d.ring.p;
    }
    ls.window = window;
    ls.block = block;
    ls.grid = grid;
    ls.smem = smem;
    return 0;
}

static int launch_ens(DevCtx &d, const void *kern, const EnsArgs &A0, const LaunchShape &ls)
{
    EnsArgs A = A0;
    A.window = ls.window;
    A.ring = ls.ring;
    A.vgrid = ls.grid;
    void *args[] = {(void *)&A};
    CK(cudaLaunchKernel(kern, dim3((ls.grid + ls.per_warp - 1) / ls.per_warp), dim3(ls.block), args, ls.smem, d.stream));
    return 0;
}

// the ahead-of-time kernel that interprets the reducer program
static const void *interpreter_kernel(int sys, int alg)
{
    const void *k = nullptr;
    by_system(sys, [&](auto sk) {
        return by_alg(alg, [&](auto ak) {
            k = (const void *)ensemble_kernel<decltype(sk)::value, decltype(ak)::value>;
            return 0;
        });
    });
    return k;
}

extern "C" int dtwa_plan_free(dtwa_plan *plan)
{
    if (!plan) return DTWA_OK;
    for (auto &dp : plan->devs) {
        if (!dp.dc) continue;
        cudaSetDevice(dp.dc->device);
        cudaStreamSynchronize(dp.dc->stream);
        if (dp.dc->copy_stream) cudaStreamSynchronize(dp.dc->copy_stream);
    }
    if (plan->ctx) plan->ctx->plan_live = false;   // (the arenas stay in the context's grow-only buffers)
    delete plan;
    return DTWA_OK;
}

extern "C" int dtwa_ensemble_launch(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const dtwa_sampler *smp,
                                    int64_t N, const double *ts, int32_t nt, const dtwa_reducer *red, int32_t nred,
                                    int32_t tolerate_errors, dtwa_plan **plan_out)
{
    if (!ctx || !plan_out) return set_err(DTWA_ERR_INVALID, "NULL context/plan");
    *plan_out = nullptr;
    TRY(check_system(sys));
    TRY(check_solver(sol, /*fixed_needs_dt=*/false));
    // N == 0 is the empty block of a rank in a one-process-per-GPU job (more ranks than trajectories, or a short last
    // chunk): zeroed arenas, no kernel -- so that every rank issues the same collectives (the shim rejects a total N < 1)
    if (N < 0) return set_err(DTWA_ERR_INVALID, "N must be >= 0");
    if (!red || nred < 1) return set_err(DTWA_ERR_INVALID, "at least one reducer is required");
    if (ctx->plan_live) return set_err(DTWA_ERR_INVALID, "the previous dtwa_plan of this context has not been freed (one ensemble in flight per context)");
    SamplerPar sp0;
    if (N == 0 && smp && (smp->kind == DTWA_SAMPLER_HOST_ARRAY || smp->kind == DTWA_SAMPLER_DEVICE_ARRAY) && !smp->u0) {
        dtwa_sampler tmp = *smp;
        static const double dummy[4] = {0, 0, 0, 0};
        tmp.u0 = dummy;
        TRY(make_sampler_par(sys, &tmp, sp0));
    } else
        TRY(make_sampler_par(sys, smp, sp0));
    std::vector<StepTab> tab;
    TRY(build_step_table(ts, nt, sol->dt, make_par(sys, sol->backwards).b, tab));
    const bool fixed = sol->alg == DTWA_ALG_RK4 || sol->alg == DTWA_ALG_RK8;
    if (fixed && ts[nt - 1] > 0 && !(sol->dt > 0)) return set_err(DTWA_ERR_INVALID, "fixed-step algorithms need dt > 0");
    const int ndev = (int)ctx->devs.size();
    if (smp->kind == DTWA_SAMPLER_DEVICE_ARRAY && ndev != 1)
        return set_err(DTWA_ERR_INVALID, "DEVICE_ARRAY sampler needs a single-device context");

    // ---- reducer program + arena layout ----
    Program prog;
    std::vector<RangeD> ranges;
    std::vector<HistDesc> hists;
    std::vector<RedLayout> reds(nred);
    int nslots = 0;
    int64_t u64_len = CNT_N + nt;
    int stage_bins = 0, stage2d_bins = 0;
    const int STAGE_BUDGET_BINS = 8192;  // x2 buffers x4 B = 64 KB of shared memory at most
    const int STAGE2D_BUDGET_BINS = 8192;  // 32 KB per CTA: keeps four CTAs of the fixed-step kernels on an SM
    auto vars = varnames_of(sys->kind);
    try {
        ExprCompiler comp(vars);
        for (int r = 0; r < nred; ++r) {
            const dtwa_reducer &R = red[r];
            RedLayout &L = reds[r];
            L.kind = R.kind;
            if (R.kind == DTWA_RED_AVERAGE) {
                if (R.nexpr < 1 || !R.exprs) return set_err(DTWA_ERR_INVALID, "AVERAGE reducer needs at least one expression");
                L.nexpr = R.nexpr;
                L.slot0 = nslots;
                for (int i = 0; i < R.nexpr; ++i) {
                    if (!R.exprs[i]) return set_err(DTWA_ERR_INVALID, "NULL expression");
                    comp.compile(R.exprs[i], prog);
                    prog.emit(OP_SUM, 0, (uint16_t)nslots++);
                }
            } else if (R.kind == DTWA_RED_SURVIVAL) {
                if (smp->kind < DTWA_SAMPLER_COHERENT_HW || smp->kind > DTWA_SAMPLER_COHERENT_HWXSU2)
                    return set_err(DTWA_ERR_INVALID, "survival probability needs a coherent_Wigner_* distribution (its pdf is the observable)");
                L.nexpr = 1;
                L.slot0 = nslots;
                prog.emit(OP_LDPDF);
                prog.emit(OP_SUM, 0, (uint16_t)nslots++);
            } else if (R.kind == DTWA_RED_HIST) {
                const bool xt = R.x_axis == DTWA_AXIS_TIME, yt = R.y_axis == DTWA_AXIS_TIME;
                if (xt && yt) return set_err(DTWA_ERR_INVALID, "at most one histogram axis may be :t");
                HistDesc H;
                std::memset(&H, 0, sizeof(H));
                H.staged_off = -1;
                H.base = u64_len;
                if (xt || yt) {
                    const char *e = xt ? R.y_expr : R.x_expr;
                    if (!e) return set_err(DTWA_ERR_INVALID, "histogram needs an expression for its non-time axis");
                    RangeD rd;
                    TRY(make_range(xt ? R.ys : R.xs, rd));
                    if (rd.len > 0x7fffffff) return set_err(DTWA_ERR_INVALID, "histogram axis too long");
                    comp.compile(e, prog);
                    prog.emit(OP_BIN, 0, (uint16_t)ranges.size());
                    ranges.push_back(rd);
                    H.two_d = 0;
                    H.n0 = (int32_t)rd.len;
                    H.n1 = 1;
                    H.tstride = rd.len;
                    L.counts_len = (int64_t)nt * rd.len;
                    // (adaptive kernels keep a long ring of output times per lane and bin straight into the arena)
                    if (fixed && stage_bins + H.n0 <= STAGE_BUDGET_BINS) {
                        H.staged_off = stage_bins;
                        stage_bins += H.n0;
                    }
                } else {
                    if (!R.x_expr || !R.y_expr) return set_err(DTWA_ERR_INVALID, "histogram needs x and y expressions");
                    RangeD rx, ry;
                    TRY(make_range(R.xs, rx));
                    TRY(make_range(R.ys, ry));
                    if (rx.len * ry.len > 0x7fffffffll) return set_err(DTWA_ERR_INVALID, "histogram matrix too large");
                    comp.compile(R.x_expr, prog);
                    prog.emit(OP_BIN, 0, (uint16_t)ranges.size());
                    ranges.push_back(rx);
                    comp.compile(R.y_expr, prog);
                    prog.emit(OP_BIN, 1, (uint16_t)ranges.size());
                    ranges.push_back(ry);
                    H.two_d = 1;
                    H.n0 = (int32_t)rx.len;
                    H.n1 = (int32_t)ry.len;
                    H.tstride = R.animate ? rx.len * ry.len : 0;
                    // a matrix shared by all times that fits the budget is accumulated in a per-CTA shared-memory tile
                    // (fixed-step kernels; the adaptive ones run one warp per CTA, 12-16 CTAs per SM: no room)
                    if (fixed && !R.animate && stage2d_bins + rx.len * ry.len <= STAGE2D_BUDGET_BINS && std::getenv("DTWA_NO_STAGE2D") == nullptr) {
                        H.staged_off = stage2d_bins;
                        stage2d_bins += (int)(rx.len * ry.len);
                    }
                    L.counts_len = (R.animate ? (int64_t)nt : 1) * rx.len * ry.len;
                }
                L.u64_off = u64_len;
                u64_len += L.counts_len;
                prog.emit(OP_HIST, 0, (uint16_t)hists.size());
                hists.push_back(H);
            } else {
                return set_err(DTWA_ERR_INVALID, "unknown reducer kind");
            }
        }
    } catch (const std::exception &e) {
        return set_err(DTWA_ERR_EXPR, e.what());
    }
    prog.emit(OP_END);
    if (prog.max_depth > VM_STACK) return set_err(DTWA_ERR_EXPR, "expression too deeply nested for the VM stack");
    if ((int)prog.ins.size() > VM_MAX_INS || (int)prog.consts.size() > VM_MAX_CONST || (int)ranges.size() > VM_MAX_RANGE ||
        (int)hists.size() > VM_MAX_HIST || nslots > VM_MAX_SLOTS)
        return set_err(DTWA_ERR_CAPACITY, "reducer set exceeds the VM limits (instructions/constants/ranges/histograms/slots)");
    if (prog.consts.empty()) prog.consts.push_back(0.0);

    // ---- device-memory bound: the count arena (animate: nt x ny x nx u64) must fit next to the other buffers ----
    for (int di = 0; di < ndev; ++di) {
        DevCtx &d = ctx->devs[di];
        CK(cudaSetDevice(d.device));
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const double need = 8.0 * (double)u64_len * 1.125 + 256.0;              // DevBuf::ensure over-allocates by 1/8
        const double avail = (double)free_b + (double)d.arena_u64.cap - 512.0 * 1024 * 1024;   // keep 512 MiB for everything else
        if ((size_t)u64_len * 8 > d.arena_u64.cap && need > avail) {
            char msg[512];
            std::snprintf(msg, sizeof(msg),
                          "the histogram counts of this call need %.2f GB of device memory (%lld u64 cells; animate=true keeps one ny x nx "
                          "matrix per output time) but device %d has %.2f GB free: use fewer output times or coarser bins, or split ts "
                          "into several calls",
                          8.0 * (double)u64_len / 1e9, (long long)u64_len, d.device, (double)free_b / 1e9);
            return set_err(DTWA_ERR_CAPACITY, msg);
        }
    }

    // ---- which kernel: the interpreter, or a run-time specialisation of it for this reducer set ----
    if (sol->chart == DTWA_CHART_BLOCH && sys->kind != DTWA_SYS_DICKE)
        return set_err(DTWA_ERR_INVALID, "chart = DTWA_CHART_BLOCH is defined for the Dicke model (the LMG field is polynomial in (Q,P) already)");
    const void *kern = interpreter_kernel(chart_sys(sys, sol), sol->alg);
    // auto mode: specialise when the estimated work repays ~1 s of NVRTC.  Fixed step: the reducers run N * nt times.
    // Adaptive: lanes reach their output times out of step, so the reducer code runs (partially masked) in most
    // iterations of the step loop -- count attempted steps instead, ~8 per unit time as a floor.
    double jit_work = (double)N * (double)nt;
    if (!fixed) jit_work = std::max(jit_work, (double)N * 8.0 * ts[nt - 1]);
    const bool want_jit = ctx->jit_mode == 1 || (ctx->jit_mode == 2 && jit_work >= 2e9);
    bool used_jit = want_jit;
    if (!want_jit && ctx->jit_mode == 2 && std::getenv("DTWA_NO_BACKGROUND_JIT") == nullptr) {
        // small call: never wait for NVRTC, but let a background thread specialise this reducer set; repeated calls
        // (parameter scans) pick the specialised kernel up once it is ready -- same bits as the interpreter
        cudaKernel_t jk;
        CK(cudaSetDevice(ctx->devs[0].device));
        const int wflag = jit_flags(sys, sol);
        if (jit_cache().get_async(chart_sys(sys, sol), sol->alg, wflag, prog, &jk) == 0) {
            kern = (const void *)jk;
            used_jit = true;
        }
    }
    if (want_jit) {
        cudaKernel_t jk;
        std::string jerr;
        CK(cudaSetDevice(ctx->devs[0].device));
        const int wflag = jit_flags(sys, sol);
        if (jit_cache().get(chart_sys(sys, sol), sol->alg, wflag, prog, &jk, jerr)) {
            // forced specialisation fails loudly; in auto mode the (slower) interpreter kernel runs instead -- still on
            // the GPU -- and the result reports jit = 0 (e.g. libnvrtc is not installed next to the driver)
            if (ctx->jit_mode == 1) return set_err(DTWA_ERR_JIT, jerr);
            set_err(DTWA_ERR_JIT, jerr);
            used_jit = false;
        } else
            kern = (const void *)jk;
    }

    dtwa_plan *plan = new dtwa_plan;
    plan->ctx = ctx;
    plan->jit = used_jit;
    plan->nt = nt;
    plan->nslots = nslots;
    plan->n_f64 = (int64_t)nt * std::max(nslots, 1);
    plan->n_u64 = u64_len;
    plan->reds = reds;
    plan->devs.resize(ndev);
    const int nv = nvar_of(sys->kind);

    // meta blob: [consts | ranges | hists | tab | ts | prog]
    auto al = [](size_t x) { return (x + 31) & ~size_t(31); };
    const size_t o_c = 0, o_r = al(o_c + sizeof(double) * prog.consts.size()), o_h = al(o_r + sizeof(RangeD) * ranges.size()),
                 o_t = al(o_h + sizeof(HistDesc) * hists.size()), o_s = al(o_t + sizeof(StepTab) * nt),
                 o_p = al(o_s + sizeof(double) * nt), meta_b = al(o_p + sizeof(VmIns) * prog.ins.size());
    std::vector<unsigned char> blob(meta_b, 0);
    std::memcpy(blob.data() + o_c, prog.consts.data(), sizeof(double) * prog.consts.size());
    if (!ranges.empty()) std::memcpy(blob.data() + o_r, ranges.data(), sizeof(RangeD) * ranges.size());
    if (!hists.empty()) std::memcpy(blob.data() + o_h, hists.data(), sizeof(HistDesc) * hists.size());
    std::memcpy(blob.data() + o_t, tab.data(), sizeof(StepTab) * nt);
    std::memcpy(blob.data() + o_s, ts, sizeof(double) * nt);
    std::memcpy(blob.data() + o_p, prog.ins.data(), sizeof(VmIns) * prog.ins.size());

    // balanced contiguous blocks: N / ndev trajectories each, the first N % ndev devices one more
    const long long blk = N / ndev, extra = N % ndev;
    int rc = 0;
    ctx->plan_live = true;
    std::vector<EnsArgs> args(ndev);
    std::vector<LaunchShape> shapes(ndev);
    std::vector<unsigned *> fail_counts(ndev, nullptr);
    std::vector<unsigned> fail_caps(ndev, 0);

    // ---- phase 1: main pass on every device (asynchronous) ----
    for (int di = 0; di < ndev && !rc; ++di) {
        DevCtx &d = ctx->devs[di];
        DevPlan &dp = plan->devs[di];
        dp.dc = &d;
        dp.lo = blk * di + std::min<long long>(di, extra);
        dp.n = blk + (di < extra ? 1 : 0);
        auto body = [&]() -> int {
            CK(cudaSetDevice(d.device));
            CK(cudaEventRecord(d.ev[0], d.stream));
            TRY(d.arena_f64.ensure(sizeof(double) * plan->n_f64));
            TRY(d.arena_u64.ensure(sizeof(unsigned long long) * plan->n_u64));
            dp.d_f64 = (double *)d.arena_f64.p;
            dp.d_u64 = (unsigned long long *)d.arena_u64.p;
            CK(cudaMemsetAsync(dp.d_f64, 0, sizeof(double) * plan->n_f64, d.stream));
            CK(cudaMemsetAsync(dp.d_u64, 0, sizeof(unsigned long long) * plan->n_u64, d.stream));
            if (dp.n == 0) {
                CK(cudaEventRecord(d.ev[1], d.stream));
                CK(cudaEventRecord(d.ev[2], d.stream));
                return 0;
            }
            TRY(d.meta.ensure(meta_b));
            CK(cudaMemcpyAsync(d.meta.p, blob.data(), meta_b, cudaMemcpyHostToDevice, d.stream));
            unsigned char *m = (unsigned char *)d.meta.p;
            EnsArgs &A = args[di];
            std::memset(&A, 0, sizeof(A));
            A.S = make_solver_args(sys, sol, (const StepTab *)(m + o_t), (const double *)(m + o_s), ts, nt, tab);
            A.smp = sp0;
            A.smp.offset = smp->offset + (uint64_t)dp.lo;
            const bool overlap_upload = smp->kind == DTWA_SAMPLER_HOST_ARRAY && fixed && d.copy_stream != nullptr && std::getenv("DTWA_NO_COPY_OVERLAP") == nullptr;
            if (smp->kind == DTWA_SAMPLER_HOST_ARRAY) {
                TRY(d.u0.ensure(sizeof(double) * dp.n * nv));
                if (!overlap_upload)
                    CK(cudaMemcpyAsync(d.u0.p, smp->u0 + dp.lo * nv, sizeof(double) * dp.n * nv, cudaMemcpyHostToDevice, d.stream));
                A.smp.u0 = (const double *)d.u0.p;
            } else if (smp->kind == DTWA_SAMPLER_DEVICE_ARRAY) {
                A.smp.u0 = smp->u0;
            }
            A.N = dp.n;
            A.attempt = 0;
            A.nslots = nslots;
            A.nins = (int)prog.ins.size();
            A.ncst = (int)prog.consts.size();
            A.nrange = (int)ranges.size();
            A.nhist = (int)hists.size();
            A.stage_bins = stage_bins;
            A.stage2d_bins = stage2d_bins;
            A.redo = nullptr;
            const bool can_redo = tolerate_errors && smp->kind != DTWA_SAMPLER_HOST_ARRAY && smp->kind != DTWA_SAMPLER_DEVICE_ARRAY;
            fail_caps[di] = (unsigned)std::min<long long>(dp.n, 1ll << 22);
            TRY(d.fail[0].ensure(sizeof(FailRec) * fail_caps[di] + 16));
            TRY(d.fail[1].ensure(sizeof(FailRec) * fail_caps[di] + 16));
            // fail buffers: [count (16 B) | records]
            CK(cudaMemsetAsync(d.fail[0].p, 0, 16, d.stream));
            A.fail_count = (unsigned *)d.fail[0].p;
            A.fail_out = can_redo ? (FailRec *)((unsigned char *)d.fail[0].p + 16) : nullptr;
            A.fail_cap = fail_caps[di];
            A.prog = (const VmIns *)(m + o_p);
            A.consts = (const double *)(m + o_c);
            A.ranges = (const RangeD *)(m + o_r);
            A.hists = (const HistDesc *)(m + o_h);
            A.u64 = dp.d_u64;
            TRY(shape_for(ctx, d, kern, sol->alg, A, dp.n, nv, shapes[di], used_jit && pair_streams(sol) ? 2 : 1,
                          used_jit && !fixed && !pair_streams(sol) ? adapt_warps_env() : 1, fixed ? 256 : adaptive_ring_cap(sys, sol, ts, nt)));
            dp.grid = shapes[di].grid;
            dp.block = shapes[di].block;
            const size_t pbytes = sizeof(double) * (size_t)dp.grid * nt * std::max(nslots, 1);
            TRY(d.partial.ensure(pbytes));
            CK(cudaMemsetAsync(d.partial.p, 0, pbytes, d.stream));
            A.partial = (double *)d.partial.p;
            // Host-supplied initial conditions, fixed-step kernels: the upload is overlapped with the integration.  The kernel deals
            // trajectory j to CTA (j / block) % grid, so the first grid * block trajectories are exactly the first batch of every
            // CTA: they are uploaded and launched on their own (LAUNCH A) while a second stream uploads the rest; LAUNCH B takes the
            // remaining batches once that copy has landed.  Every trajectory meets the same thread, in the same order, accumulating
            // into the same per-CTA rows as in one launch: bit-identical results; the cost is one more launch and one CTA-level
            // synchronisation after batches of equal length.  (The adaptive kernels are not split: their lanes finish a pass at
            // different times.)
            const long long first_pass = (long long)shapes[di].grid * shapes[di].block;
            if (overlap_upload && dp.n >= 8 * first_pass) {
                const double *h0 = smp->u0 + dp.lo * nv;
                CK(cudaMemcpyAsync(d.u0.p, h0, sizeof(double) * first_pass * nv, cudaMemcpyHostToDevice, d.stream));
                CK(cudaEventRecord(d.ev_copy_go, d.stream));   // (orders the second stream behind every earlier user of d.u0)
                CK(cudaEventRecord(d.ev[1], d.stream));
                EnsArgs A1 = A;
                A1.N = first_pass;
                TRY(launch_ens(d, kern, A1, shapes[di]));
                plan->launches++;
                CK(cudaStreamWaitEvent(d.copy_stream, d.ev_copy_go, 0));
                CK(cudaMemcpyAsync((double *)d.u0.p + first_pass * nv, h0 + first_pass * nv, sizeof(double) * (dp.n - first_pass) * nv, cudaMemcpyHostToDevice,
                                   d.copy_stream));
                CK(cudaEventRecord(d.ev_copy_done, d.copy_stream));
                CK(cudaStreamWaitEvent(d.stream, d.ev_copy_done, 0));
                EnsArgs A2 = A;
                A2.N = dp.n - first_pass;
                A2.smp.u0 = A.smp.u0 + first_pass * nv;
                A2.smp.offset = A.smp.offset + (uint64_t)first_pass;
                TRY(launch_ens(d, kern, A2, shapes[di]));
                plan->launches++;
                return 0;
            }
            if (overlap_upload) CK(cudaMemcpyAsync(d.u0.p, smp->u0 + dp.lo * nv, sizeof(double) * dp.n * nv, cudaMemcpyHostToDevice, d.stream));
            CK(cudaEventRecord(d.ev[1], d.stream));
            TRY(launch_ens(d, kern, A, shapes[di]));
            plan->launches++;
            return 0;
        };
        rc = body();
    }

    // ---- phase 2: resampling rounds (src/TWA.jl:186-209), interleaved over the devices: every device's failure count
    //      is requested before any stream is waited for, so device i+1 is not idle while device i resamples ----
    {
        std::vector<int> cur(ndev, 0);
        std::vector<char> active(ndev, 0);
        std::vector<unsigned> cnt(ndev, 0);
        std::vector<std::vector<FailRec>> recs(ndev);
        bool any_active = false;
        for (int di = 0; di < ndev; ++di) active[di] = plan->devs[di].n > 0, any_active |= (bool)active[di];
        auto on = [&](int di, auto &&f) -> int {   // run f on device di, folding its status into rc
            if (rc || !active[di]) return 0;
            if (cudaSetDevice(ctx->devs[di].device) != cudaSuccess) return rc = set_err(DTWA_ERR_CUDA, "cudaSetDevice failed");
            return rc = f(ctx->devs[di], plan->devs[di], args[di]);
        };
        for (int round = 1; !rc && any_active; ++round) {
            for (int di = 0; di < ndev; ++di)
                on(di, [&](DevCtx &d, DevPlan &, EnsArgs &) -> int {
                    CK(cudaMemcpyAsync(&cnt[di], d.fail[cur[di]].p, sizeof(unsigned), cudaMemcpyDeviceToHost, d.stream));
                    return 0;
                });
            for (int di = 0; di < ndev; ++di)
                on(di, [&](DevCtx &d, DevPlan &dp, EnsArgs &A) -> int {
                    CK(cudaStreamSynchronize(d.stream));
                    if (!A.fail_out) {
                        active[di] = 0;
                        if (!tolerate_errors) {  // the reference rethrows the first error (src/TWA.jl:199-201)
                            unsigned long long nfail = 0;
                            CK(cudaMemcpyAsync(&nfail, dp.d_u64 + CNT_FAILED, sizeof(nfail), cudaMemcpyDeviceToHost, d.stream));
                            CK(cudaStreamSynchronize(d.stream));
                            if (nfail)
                                return set_err(DTWA_ERR_BOUNDS, "a trajectory left the domain (DomainError in the reference) and tolerate_errors=false");
                        }
                        return 0;
                    }
                    if (cnt[di] == 0) {
                        active[di] = 0;
                        return 0;
                    }
                    if (cnt[di] > fail_caps[di]) cnt[di] = fail_caps[di];
                    if (round > 100)
                        return set_err(DTWA_ERR_TOO_MANY_FAILURES, "Too many errors: a trajectory failed in more than 100 consecutive resampling rounds");
                    recs[di].resize(cnt[di]);
                    CK(cudaMemcpyAsync(recs[di].data(), (unsigned char *)d.fail[cur[di]].p + 16, sizeof(FailRec) * cnt[di], cudaMemcpyDeviceToHost, d.stream));
                    return 0;
                });
            any_active = false;
            for (int di = 0; di < ndev; ++di)
                on(di, [&](DevCtx &d, DevPlan &, EnsArgs &A) -> int {
                    CK(cudaStreamSynchronize(d.stream));   // (also: the previous round's upload from recs[di] has completed)
                    // sort the failure list by trajectory index so that the round is reproducible
                    std::sort(recs[di].begin(), recs[di].end(), [](const FailRec &x, const FailRec &y) { return x.idx < y.idx; });
                    const int c = cur[di], nxt = c ^ 1;
                    CK(cudaMemcpyAsync((unsigned char *)d.fail[c].p + 16, recs[di].data(), sizeof(FailRec) * cnt[di], cudaMemcpyHostToDevice, d.stream));
                    CK(cudaMemsetAsync(d.fail[nxt].p, 0, 16, d.stream));
                    A.redo = (const FailRec *)((unsigned char *)d.fail[c].p + 16);
                    A.fail_count = (unsigned *)d.fail[nxt].p;
                    A.fail_out = (FailRec *)((unsigned char *)d.fail[nxt].p + 16);
                    A.N = cnt[di];
                    A.attempt = (uint32_t)round;
                    LaunchShape ls = shapes[di];
                    const int per_unit = fixed ? ls.block : 32;   // trajectories of the first pass of a CTA (fixed step) / warp stream (adaptive)
                    ls.grid = (int)std::max(1ll, std::min<long long>((cnt[di] + per_unit - 1) / per_unit, shapes[di].grid));
                    TRY(launch_ens(d, kern, A, ls));
                    plan->launches++;
                    plan->rounds = std::max(plan->rounds, round);
                    cur[di] = nxt;
                    any_active = true;
                    return 0;
                });
        }
        // fixed-order merge of the per-CTA partial sums
        for (int di = 0; di < ndev && !rc; ++di) {
            DevCtx &d = ctx->devs[di];
            DevPlan &dp = plan->devs[di];
            if (dp.n == 0) continue;
            auto body = [&]() -> int {
                CK(cudaSetDevice(d.device));
                if (nslots > 0) {
                    const long long nsum = (long long)nt * nslots;
                    finalize_sums_kernel<<<blocks_for(nsum, 256), 256, 0, d.stream>>>((const double *)d.partial.p, dp.grid, nsum, dp.d_f64);
                    CK(cudaGetLastError());
                    plan->launches++;
                }
                CK(cudaEventRecord(d.ev[2], d.stream));
                return 0;
            };
            rc = body();
        }
    }

    // ---- phase 3: one grouped all-reduce over NVLink when the context spans several devices ----
    if (!rc && ndev > 1) {
        NcclApi &n = ctx->nccl;
        int r = n.GroupStart();
        for (int di = 0; di < ndev && !r; ++di) {
            DevPlan &dp = plan->devs[di];
            r = n.AllReduce(dp.d_f64, dp.d_f64, (size_t)plan->n_f64, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->comms[di], dp.dc->stream);
            if (!r)
                r = n.AllReduce(dp.d_u64, dp.d_u64, (size_t)plan->n_u64, /*ncclUint64*/ 5, /*ncclSum*/ 0, ctx->comms[di], dp.dc->stream);
        }
        int r2 = n.GroupEnd();
        if (r || r2) rc = set_err(DTWA_ERR_NCCL, std::string("ncclAllReduce: ") + n.GetErrorString(r ? r : r2));
    }
    // every kernel (incl. the merge of the partial sums) has completed when launch returns, so a caller
    // may all-reduce the arenas on any stream of its own before dtwa_plan_fetch
    for (int di = 0; di < ndev && !rc; ++di) {
        if (cudaSetDevice(ctx->devs[di].device) != cudaSuccess || cudaStreamSynchronize(ctx->devs[di].stream) != cudaSuccess)
            rc = set_err(DTWA_ERR_CUDA, std::string("stream synchronize: ") + cudaGetErrorString(cudaGetLastError()));
    }
    if (rc) {
        dtwa_plan_free(plan);
        return rc;
    }
    *plan_out = plan;
    return DTWA_OK;
}

extern "C" int dtwa_plan_arenas(dtwa_plan *plan, void **d_f64, int64_t *n_f64, void **d_u64, int64_t *n_u64)
{
    if (!plan) return set_err(DTWA_ERR_INVALID, "NULL plan");
    if (d_f64) *d_f64 = plan->devs[0].d_f64;
    if (n_f64) *n_f64 = plan->n_f64;
    if (d_u64) *d_u64 = plan->devs[0].d_u64;
    if (n_u64) *n_u64 = plan->n_u64;
    return DTWA_OK;
}

extern "C" int dtwa_plan_fetch(dtwa_plan *plan, dtwa_result *res)
{
    if (!plan || !res) return set_err(DTWA_ERR_INVALID, "NULL plan/result");
    DevPlan &dp = plan->devs[0];
    DevCtx &d = *dp.dc;
    CK(cudaSetDevice(d.device));
    // validate capacities before copying anything
    for (size_t r = 0; r < plan->reds.size(); ++r) {
        const RedLayout &L = plan->reds[r];
        if (!res->red) return set_err(DTWA_ERR_INVALID, "result.red is NULL");
        const dtwa_reducer_out &o = res->red[r];
        if (L.kind == DTWA_RED_HIST) {
            if (!o.counts || o.counts_len < L.counts_len) return set_err(DTWA_ERR_CAPACITY, "counts buffer too small");
        } else {
            if (!o.sums || o.sums_len < (int64_t)plan->nt * L.nexpr) return set_err(DTWA_ERR_CAPACITY, "sums buffer too small");
        }
    }
    std::vector<double> f64(plan->n_f64);
    std::vector<unsigned long long> head(CNT_N + plan->nt);
    CK(cudaMemcpyAsync(f64.data(), dp.d_f64, sizeof(double) * plan->n_f64, cudaMemcpyDeviceToHost, d.stream));
    CK(cudaMemcpyAsync(head.data(), dp.d_u64, sizeof(unsigned long long) * head.size(), cudaMemcpyDeviceToHost, d.stream));
    for (size_t r = 0; r < plan->reds.size(); ++r) {
        const RedLayout &L = plan->reds[r];
        if (L.kind == DTWA_RED_HIST)
            CK(cudaMemcpyAsync(res->red[r].counts, dp.d_u64 + L.u64_off, sizeof(unsigned long long) * L.counts_len, cudaMemcpyDeviceToHost, d.stream));
    }
    CK(cudaEventRecord(d.ev[3], d.stream));
    CK(cudaStreamSynchronize(d.stream));
    for (size_t r = 0; r < plan->reds.size(); ++r) {
        const RedLayout &L = plan->reds[r];
        if (L.kind == DTWA_RED_HIST) continue;
        double *o = res->red[r].sums;
        for (int k = 0; k < plan->nt; ++k)
            for (int i = 0; i < L.nexpr; ++i) o[(size_t)k * L.nexpr + i] = f64[(size_t)k * plan->nslots + L.slot0 + i];
    }
    if (res->n_valid)
        for (int k = 0; k < plan->nt; ++k) res->n_valid[k] = head[CNT_N + k];
    res->n_failed = head[CNT_FAILED];
    res->n_unfinished = head[CNT_UNFINISHED];
    res->steps_accepted = head[CNT_ACC];
    res->steps_rejected = head[CNT_REJ];
    res->lane_steps = head[CNT_LANE];
    res->resample_rounds = plan->rounds;
    res->jit = plan->jit ? 1 : 0;
    res->launches = plan->launches;
    float tms = 0.f;
    double kmax = 0.0;
    for (auto &p : plan->devs) {  // max over devices, each timed on its own stream
        if (!p.dc) continue;
        cudaSetDevice(p.dc->device);
        cudaEventSynchronize(p.dc->ev[2]);
        float m = 0.f;
        if (cudaEventElapsedTime(&m, p.dc->ev[1], p.dc->ev[2]) == cudaSuccess) kmax = std::max(kmax, (double)m);
    }
    cudaSetDevice(d.device);
    cudaEventElapsedTime(&tms, d.ev[0], d.ev[3]);
    cudaGetLastError();
    res->kernel_ms = kmax;
    res->total_ms = tms;
    plan->fetched = true;
    return DTWA_OK;
}

extern "C" int dtwa_ensemble(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const dtwa_sampler *smp, int64_t N,
                             const double *ts, int32_t nt, const dtwa_reducer *red, int32_t nred, int32_t tolerate_errors,
                             dtwa_result *res)
{
    if (!res) return set_err(DTWA_ERR_INVALID, "result is NULL");
    dtwa_plan *plan = nullptr;
    TRY(dtwa_ensemble_launch(ctx, sys, sol, smp, N, ts, nt, red, nred, tolerate_errors, &plan));
    int rc = dtwa_plan_fetch(plan, res);
    dtwa_plan_free(plan);
    return rc;
}

// ------------------------------------------------------------------------------------------------
extern "C" int dtwa_fp64_peak(dtwa_ctx *ctx, double *tflops, double *ms_out)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    const int block = 256, grid = d.sm_count * 8, iters = 4096;
    TRY(d.scratch_a.ensure(sizeof(double) * (size_t)grid * block));
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(d.ev[0], d.stream));
        fp64_peak_kernel<<<grid, block, 0, d.stream>>>((double *)d.scratch_a.p, iters, 0.999999, 1e-7);
        CK(cudaGetLastError());
        CK(cudaEventRecord(d.ev[3], d.stream));
        CK(cudaStreamSynchronize(d.stream));
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, d.ev[0], d.ev[3]));
        if (rep > 0 && m < best) best = m;
    }
    const double flops = (double)grid * block * (double)iters * 64.0 * 2.0;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return DTWA_OK;
}

// K0b: DFMA rate with three distinct register operands per instruction (see fp64_peak_regs_kernel)
extern "C" int dtwa_fp64_peak_regs(dtwa_ctx *ctx, double *tflops, double *ms_out)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    const int block = 256, grid = d.sm_count * 8, iters = 4096;
    TRY(d.scratch_a.ensure(sizeof(double) * (size_t)grid * block));
    TRY(d.scratch_b.ensure(sizeof(double) * 64));
    double host[64];
    for (int i = 0; i < 64; ++i) host[i] = 1.0 + 1e-3 * i;
    CK(cudaMemcpyAsync(d.scratch_b.p, host, sizeof(host), cudaMemcpyHostToDevice, d.stream));
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(d.ev[0], d.stream));
        fp64_peak_regs_kernel<<<grid, block, 0, d.stream>>>((double *)d.scratch_a.p, (const double *)d.scratch_b.p, iters);
        CK(cudaGetLastError());
        CK(cudaEventRecord(d.ev[3], d.stream));
        CK(cudaStreamSynchronize(d.stream));
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, d.ev[0], d.ev[3]));
        if (rep > 0 && m < best) best = m;
    }
    const double flops = (double)grid * block * (double)iters * 64.0 * 2.0;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return DTWA_OK;
}

// K0c: DFMA rate with `chains` independent chains per thread and `warps_per_sm` one-warp CTAs per SM (chains in {1,2,4,8,16})
extern "C" int dtwa_fp64_chains(dtwa_ctx *ctx, int chains, int warps_per_sm, int nreg, double *tflops, double *ms_out)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (warps_per_sm < 1 || warps_per_sm > 32) return set_err(DTWA_ERR_INVALID, "warps_per_sm must be in [1, 32]");
    if (nreg < 1 || nreg > 3) return set_err(DTWA_ERR_INVALID, "nreg must be 1, 2 or 3");
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    const int block = 32, grid = d.sm_count * warps_per_sm, iters = 16384;
    TRY(d.scratch_a.ensure(sizeof(double) * (size_t)grid * block));
    TRY(d.scratch_b.ensure(sizeof(double) * 64));
    double host[64];
    for (int i = 0; i < 64; ++i) host[i] = 1.0 + 1e-3 * i;
    CK(cudaMemcpyAsync(d.scratch_b.p, host, sizeof(host), cudaMemcpyHostToDevice, d.stream));
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(d.ev[0], d.stream));
        double *o = (double *)d.scratch_a.p;
        const double *in = (const double *)d.scratch_b.p;
#define DTWA_CHAINS_CASE(C)                                                                                              \
    case C:                                                                                                              \
        if (nreg == 1) fp64_chains_kernel<C, 1><<<grid, block, 0, d.stream>>>(o, in, iters, 0.999999, 1e-7);              \
        else if (nreg == 2) fp64_chains_kernel<C, 2><<<grid, block, 0, d.stream>>>(o, in, iters, 0.999999, 1e-7);         \
        else fp64_chains_kernel<C, 3><<<grid, block, 0, d.stream>>>(o, in, iters, 0.999999, 1e-7);                        \
        break;
        switch (chains) {
            DTWA_CHAINS_CASE(1)
            DTWA_CHAINS_CASE(2)
            DTWA_CHAINS_CASE(4)
            DTWA_CHAINS_CASE(8)
            DTWA_CHAINS_CASE(16)
        default: return set_err(DTWA_ERR_INVALID, "chains must be 1, 2, 4, 8 or 16");
        }
#undef DTWA_CHAINS_CASE
        CK(cudaGetLastError());
        CK(cudaEventRecord(d.ev[3], d.stream));
        CK(cudaStreamSynchronize(d.stream));
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, d.ev[0], d.ev[3]));
        if (rep > 0 && m < best) best = m;
    }
    const double flops = (double)grid * block * (double)iters * 64.0 * 2.0;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return DTWA_OK;
}

// K0d: K0's DFMA chains with `alu_per_8` instructions of another pipe woven in per 8 DFMAs (0, 2, 4, 8, 16), kind 0 = integer
// add, 1 = FP32 FMA, 2 = MUFU; `warps_per_sm` one-warp CTAs per SM.  Reports the DFMA rate alone.
extern "C" int dtwa_fp64_mix(dtwa_ctx *ctx, int alu_per_8, int kind, int warps_per_sm, double *tflops, double *ms_out)
{
    if (!ctx) return set_err(DTWA_ERR_INVALID, "NULL context");
    if (warps_per_sm < 1 || warps_per_sm > 32) return set_err(DTWA_ERR_INVALID, "warps_per_sm must be in [1, 32]");
    if (kind < 0 || kind > 4) return set_err(DTWA_ERR_INVALID, "kind must be 0 .. 4");
    DevCtx &d = ctx->devs[0];
    CK(cudaSetDevice(d.device));
    const int block = 32, grid = d.sm_count * warps_per_sm, iters = 16384;
    TRY(d.scratch_a.ensure(sizeof(double) * (size_t)grid * block));
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(d.ev[0], d.stream));
        double *o = (double *)d.scratch_a.p;
#define DTWA_MIX_CASE(A)                                                                                       \
    case A:                                                                                                    \
        if (kind == 0) fp64_mix_kernel<A, 0><<<grid, block, 0, d.stream>>>(o, iters, 0.999999, 1e-7);          \
        else if (kind == 1) fp64_mix_kernel<A, 1><<<grid, block, 0, d.stream>>>(o, iters, 0.999999, 1e-7);     \
        else if (kind == 2) fp64_mix_kernel<A, 2><<<grid, block, 0, d.stream>>>(o, iters, 0.999999, 1e-7);     \
        else if (kind == 3) fp64_mix_kernel<A, 3><<<grid, block, 0, d.stream>>>(o, iters, 0.999999, 1e-7);     \
        else fp64_mix_kernel<A, 4><<<grid, block, 0, d.stream>>>(o, iters, 0.999999, 1e-7);                    \
        break;
        switch (alu_per_8) {
            DTWA_MIX_CASE(0)
            DTWA_MIX_CASE(2)
            DTWA_MIX_CASE(4)
            DTWA_MIX_CASE(8)
            DTWA_MIX_CASE(16)
        default: return set_err(DTWA_ERR_INVALID, "alu_per_8 must be 0, 2, 4, 8 or 16");
        }
#undef DTWA_MIX_CASE
        CK(cudaGetLastError());
        CK(cudaEventRecord(d.ev[3], d.stream));
        CK(cudaStreamSynchronize(d.stream));
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, d.ev[0], d.ev[3]));
        if (rep > 0 && m < best) best = m;
    }
    const double flops = (double)grid * block * (double)iters * 64.0 * 2.0;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return DTWA_OK;
}

// ------------------------------------------------------------------------------------------------
// Compile-only check of the run-time specialisation (needs libnvrtc but no GPU): generates the
// specialised source for `nexpr` averaged observables, compiles it for sm_100a and reports the
// cubin size; the generated source is copied to src_out (truncated to src_cap, always terminated).
extern "C" int dtwa_jit_compile(const dtwa_system *sys, const dtwa_solver *sol, const char *const *exprs, int32_t nexpr,
                                char *src_out, int64_t src_cap, int64_t *cubin_bytes)
{
    TRY(check_system(sys));
    if (!sol || sol->alg < DTWA_ALG_RK4 || sol->alg > DTWA_ALG_TSIT5) return set_err(DTWA_ERR_INVALID, "unknown solver alg");
    if (!exprs || nexpr < 1) return set_err(DTWA_ERR_INVALID, "at least one expression is required");
    Program prog;
    try {
        auto vars = varnames_of(sys->kind);
        ExprCompiler comp(vars);
        for (int i = 0; i < nexpr; ++i) {
            comp.compile(exprs[i], prog);
            prog.emit(OP_SUM, 0, (uint16_t)i);
        }
    } catch (const std::exception &e) {
        return set_err(DTWA_ERR_EXPR, e.what());
    }
    prog.emit(OP_END);
    const std::string src = jit_source(chart_sys(sys, sol), sol->alg, jit_flags(sys, sol), prog);
    if (src_out && src_cap > 0) {
        std::snprintf(src_out, (size_t)src_cap, "%s", src.c_str());
    }
    std::vector<char> cubin;
    std::string log;
    NvrtcApi api;
    if (jit_compile(api, src, cubin, log)) return set_err(DTWA_ERR_JIT, log);
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return DTWA_OK;
}
