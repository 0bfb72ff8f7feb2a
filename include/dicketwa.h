/*
 * dicketwa.h -- C ABI of libdicketwa.so, the B200 (sm_100a) implementation of DickeModel.jl's
 * Truncated-Wigner (TWA) ensemble hot path.
 *
 * The reference is pure Julia and has no FFI of its own; its extension points on this path are
 * Julia dispatch and closures.  Each entry point below therefore names the
 * reference *function* it replaces (paths relative to the reference checkout).  A Julia package
 * that `ccall`s these functions keeps the reference's module/function names
 * (dickemodel.jl_b200/julia/DickeModelB200.jl); the Python ctypes mirror in
 * dickemodel.jl_b200/ is the executable harness because Julia is absent from the build image.
 *
 * Conventions
 *  - plain pointers and sizes, POD structs, no C++/torch types;
 *  - the caller owns every host buffer, the library owns all device memory inside dtwa_ctx;
 *  - every function returns an int status (DTWA_OK = 0, negative = error) and never throws;
 *    dtwa_last_error() returns a thread-local message for the last failure;
 *  - calls are blocking unless stated; a context is not re-entrant, distinct contexts may be used
 *    from distinct threads;
 *  - there is NO CPU fallback: without a usable sm_100-class CUDA device dtwa_create fails.
 *  - state ordering: Dicke u = [Q,q,P,p] (src/ClassicalDicke.jl:21), LMG u = [Q,P]
 *    (src/ClassicalLMG.jl:15); all arrays row-major, trajectory index slowest.
 */
#ifndef DICKETWA_H
#define DICKETWA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTWA_VERSION 200 /* 0.2.0: dtwa_solver grew the step-controller fields */

/* ---- status codes ---- */
#define DTWA_OK 0
#define DTWA_ERR_INVALID (-1)      /* bad argument (the reference's error(...) calls) */
#define DTWA_ERR_CUDA (-2)         /* CUDA runtime failure */
#define DTWA_ERR_NO_DEVICE (-3)    /* no CUDA device: this library has no CPU path */
#define DTWA_ERR_EXPR (-4)         /* observable expression does not parse / unknown variable (src/TWA.jl:291-293) */
#define DTWA_ERR_NCCL (-5)
#define DTWA_ERR_TOO_MANY_FAILURES (-6) /* >100 consecutive resampling rounds (src/TWA.jl:202-206) */
#define DTWA_ERR_BOUNDS (-7)       /* initial condition out of bounds (src/ClassicalSystems.jl:100-102) */
#define DTWA_ERR_CAPACITY (-8)     /* caller buffer too small / internal limit exceeded */
#define DTWA_ERR_JIT (-9)          /* run-time specialisation failed (libnvrtc missing or compile error) */

/* ---- per-trajectory status (dtwa_integrate) ---- */
#define DTWA_TRAJ_OK 0
#define DTWA_TRAJ_DOMAIN 1   /* left the chart 4-P^2-Q^2>0 or became non-finite: the reference throws */
#define DTWA_TRAJ_MAXITERS 2 /* adaptive solver used up maxiters attempts */
#define DTWA_TRAJ_BAD_IC 3   /* initial condition out of bounds */
#define DTWA_LYAP_MAXITERS 5 /* dtwa_lyapunov: no convergence within maxiters intervals (the reference: error("Maximum iterations")) */

typedef struct dtwa_ctx dtwa_ctx;
typedef struct dtwa_plan dtwa_plan;

/* ---- system: replaces ClassicalDickeSystem (src/ClassicalDicke.jl:48-52) and
 *      ClassicalLMGSystem (src/ClassicalLMG.jl:32-36); params in the reference's own order:
 *      Dicke {w0, w, gamma}, LMG {Omega, xi}. ---- */
#define DTWA_SYS_DICKE 0
#define DTWA_SYS_LMG 1
typedef struct {
    int32_t kind;
    int32_t reserved;
    double params[4];
} dtwa_system;

/* ---- solver: the keyword arguments ClassicalSystems.integrate forwards to solve()
 *      (src/ClassicalSystems.jl:78-88,138).  tol is abstol=reltol; dtmin<=0 means dtmin=tol with
 *      force_dtmin (the reference default); maxiters<=0 means 100000 (OrdinaryDiffEq default).
 *      Fixed-step algorithms take ceil(interval/dt) equal sub-steps per output interval so that
 *      every ts[k] is hit exactly. ---- */
#define DTWA_ALG_RK4 0     /* classical 4th order, fixed step dt */
#define DTWA_ALG_RK8 1     /* DOP853 8th-order weights, fixed step dt */
#define DTWA_ALG_DP5 2     /* Dormand-Prince 5(4), adaptive  (reference: integrator_alg=DP5()) */
#define DTWA_ALG_DOP853 3  /* Hairer DOP853 8(5,3), adaptive (reference: integrator_alg=DP8()) */
#define DTWA_ALG_TSIT5 4   /* Tsitouras 5(4), adaptive       (reference: integrator_alg=Tsit5()) */
/* Step-size controller of the adaptive pairs.
 *  DTWA_CTRL_ODE (0, default): OrdinaryDiffEq's, i.e. what the reference's solve() runs (src/ClassicalSystems.jl:138):
 *    q = EEst^beta1 / qold^beta2 / gamma clamped to [1/qmax, 1/qmin]; accept when EEst <= 1, then dt/q and
 *    qold = max(EEst, qoldinit); rejected: dt / min(1/qmin, EEst^beta1/gamma); isoutofdomain: dt*qmin.
 *    beta1/beta2/qmin/qmax/gamma/qoldinit are solve()'s keywords of the same names.  beta1 > 0 sets beta1 and beta2
 *    together (beta2 = 0 is then taken literally); any other field <= 0 -- and a zero-initialised struct -- selects
 *    OrdinaryDiffEq's default for the algorithm: DP5 0.17/0.04/0.2/10, DP8 0.125/0/(1/3)/6, Tsit5 0.14/0.08/0.2/10, gamma 0.9,
 *    qoldinit 1e-4.
 *    The shim passes the generic order-8 gains 7/80, 1/20, 1/5, 10 for integrator_alg = TsitPap8() / Vern8().
 *  DTWA_CTRL_SCIPY (1): SciPy's rk.py controller (factor 0.9 err^(-1/(q+1)) in [0.2,10], no growth after a rejection);
 *    the golden SciPy step counts (tests/golden/scipy_pairs.json) pin the tableaux through it.
 * Only dtwa_integrate and dtwa_ensemble(_launch) honour `controller`; the variational / section entry points
 * (dtwa_integrate_fm, dtwa_lyapunov, dtwa_integrate_events*) always run DTWA_CTRL_SCIPY. */
#define DTWA_CTRL_ODE 0
#define DTWA_CTRL_SCIPY 1
/* Coordinates the Dicke trajectories are integrated in.  DTWA_CHART_QP (0, default): the reference's (Q,q,P,p)
 * (src/ClassicalDicke.jl:11-17) -- the parity contract.  DTWA_CHART_BLOCH (1, opt-in): the spin as the Cartesian unit vector
 * (jx,jy,jz) = (Q s/2, -P s/2, (Q^2+P^2)/2 - 1), in which the same equations of motion are bilinear: no square root, no
 * division, no singular passage at Q^2+P^2 -> 4.  Exact solutions coincide; numerical ones differ by the truncation error
 * of the method, so results are NOT step-for-step comparable with the reference's.  Inputs, outputs and observables stay
 * in (Q,q,P,p) either way. */
#define DTWA_CHART_QP 0
#define DTWA_CHART_BLOCH 1
/* How the adaptive algorithms reach the output times ts[k].
 *  DTWA_OUTPUT_STEP (0, default): the step is clipped so that it lands on every ts[k] (the un-clipped proposal is kept for
 *    the next step) -- states at ts[k] carry the integrator's own accuracy.
 *  DTWA_OUTPUT_HERMITE (1): the reference's semantics.  Steps ignore the output times; the state at ts[k] is OrdinaryDiffEq's
 *    3rd-order Hermite interpolant over the accepted step that contains it -- what FunctionCallingCallback(funcat = ts)
 *    evaluates for the reference default TsitPap8 (src/TWA.jl:180).  Fewer steps on dense output grids (ts = 0:0.05:40 at
 *    tol 1e-12: 1.3-1.5x, at 1e-8: 2-2.8x), interpolation error O(h^4) at the outputs.  A step is still clipped at the last
 *    output time and so that it never spans more than 32 output times. */
#define DTWA_OUTPUT_STEP 0
#define DTWA_OUTPUT_HERMITE 1
typedef struct {
    int32_t alg;
    int32_t backwards; /* the reference's `integate_backwards` [sic] */
    double dt;
    double tol;
    double dtmin;
    int64_t maxiters;
    int32_t controller; /* DTWA_CTRL_* */
    int32_t chart;      /* DTWA_CHART_* (dtwa_integrate and dtwa_ensemble(_launch) only) */
    double beta1, beta2, qmin, qmax, gamma, qoldinit;
    int32_t output;     /* DTWA_OUTPUT_* (adaptive algorithms in dtwa_integrate and dtwa_ensemble(_launch) only) */
    int32_t reserved;
} dtwa_solver;

/* ---- sampler: replaces PhaseSpaceDistribution.sample of coherent_Wigner_HW / _SU2 / _HWxSU2
 *      (src/TWA.jl:689-764).  center = {q0,p0} | {Q0,P0} | {Q0,q0,P0,p0}; hbar = 1/j.
 *      Counter-based Philox4x32-10: key = seed, counter = (offset + i, attempt, block), so the
 *      sample set does not depend on the launch shape or on how many GPUs share the ensemble.
 *      HOST_ARRAY / DEVICE_ARRAY take caller-supplied initial conditions u0[N][nvar]. ---- */
#define DTWA_SAMPLER_HOST_ARRAY 0
#define DTWA_SAMPLER_COHERENT_HW 1
#define DTWA_SAMPLER_COHERENT_SU2 2
#define DTWA_SAMPLER_COHERENT_HWXSU2 3
#define DTWA_SAMPLER_DEVICE_ARRAY 4
typedef struct {
    int32_t kind;
    int32_t reserved;
    double center[4];
    double hbar;
    uint64_t seed;
    uint64_t offset;  /* global index of this call's first trajectory */
    const double *u0; /* HOST_ARRAY: host pointer, DEVICE_ARRAY: device pointer; else NULL */
} dtwa_sampler;

/* ---- reducers: replace the MonteCarloSystem closures (src/TWA.jl:37-51) built by
 *      mcs_for_averaging (:336-370), mcs_for_distributions (:413-543) and
 *      mcs_for_survival_probability (:652-668).  Expressions are UTF-8 strings over the system's
 *      varnames with + - * / ^ ( ) unary minus, numbers, pi, and
 *      sqrt sin cos tan exp log abs acos asin atan min max pow atan2.  Several reducers in one
 *      call share the same trajectories (mcs_chain, src/TWA.jl:235-274). ---- */
#define DTWA_RED_AVERAGE 0
#define DTWA_RED_HIST 1
#define DTWA_RED_SURVIVAL 2
#define DTWA_AXIS_EXPR 0
#define DTWA_AXIS_TIME 1
typedef struct {
    double start, stop; /* LinRange(start, stop, len), both ends inclusive (src/TWA.jl:53-64) */
    int64_t len;
} dtwa_range;
typedef struct {
    int32_t kind;
    int32_t nexpr;            /* AVERAGE: number of observables */
    const char *const *exprs; /* AVERAGE: nexpr strings */
    int32_t x_axis, y_axis;   /* HIST: DTWA_AXIS_* ; at most one may be TIME */
    const char *x_expr;       /* HIST: used when x_axis == EXPR */
    const char *y_expr;
    dtwa_range xs, ys;        /* HIST: bins of the EXPR axes */
    int32_t animate;          /* HIST with two EXPR axes: one matrix per time (src/TWA.jl:480-489) */
    int32_t reserved;
} dtwa_reducer;

/* Result buffers, caller-allocated.  Raw (un-normalised) accumulators: the shim divides by N
 * (src/TWA.jl:362,518,664).
 *   AVERAGE : sums[nt][nexpr]
 *   SURVIVAL: sums[nt]
 *   HIST    : x TIME -> counts[nt][ys.len]; y TIME -> counts[nt][xs.len];
 *             two EXPR axes -> counts[ys.len][xs.len] (all times) or counts[nt][ys.len][xs.len] (animate) */
typedef struct {
    double *sums;
    int64_t sums_len; /* capacity in elements, checked */
    uint64_t *counts;
    int64_t counts_len;
} dtwa_reducer_out;

typedef struct {
    dtwa_reducer_out *red;   /* [nred] */
    uint64_t *n_valid;       /* [nt] trajectories that contributed at each time (may be NULL) */
    uint64_t n_failed;       /* trajectories that hit DOMAIN (each resampling attempt counts) */
    uint64_t n_unfinished;   /* trajectories that ended with MAXITERS or could not be resampled */
    uint64_t steps_accepted; /* trajectory-steps taken */
    uint64_t steps_rejected; /* adaptive only */
    uint64_t lane_steps;     /* 32 x warp-level step iterations: steps/(lane_steps) = lane efficiency */
    int32_t resample_rounds;
    int32_t launches;        /* kernels launched by this call */
    int32_t jit;             /* 1 when the run-time specialised kernel ran, 0 for the interpreter */
    int32_t reserved;
    double kernel_ms;        /* CUDA-event time of the ensemble kernels on the context's stream */
    double total_ms;         /* CUDA-event time incl. host<->device copies */
} dtwa_result;

/* ---- context ---- */
const char *dtwa_last_error(void);
int dtwa_version(void);
/* device_ids[ndev]; ndev>1 shards [0,N) by contiguous index blocks over the devices of one
 * process and merges the per-time arrays with one grouped ncclAllReduce (libnccl is dlopen'ed).
 * stream: an existing cudaStream_t (as void*) to enqueue on when ndev==1, or NULL for a private
 * non-blocking one (pass cudaStreamLegacy / cudaStreamPerThread to name the default streams).
 * The batch entry points on independent items (dtwa_integrate, _integrate_fm, _lyapunov, _integrate_events(_packed),
 * _shell_quadrature) need no exchange at all: with ndev>1 every device takes one contiguous block of the items, the
 * blocks run concurrently (one host thread per device) and the caller's arrays are filled block by block -- the same
 * bits as on one device.  The remaining single-item helpers (dtwa_sample, _pdf, _hamiltonian, ...) use device_ids[0]. */
int dtwa_create(dtwa_ctx **out, const int *device_ids, int ndev, void *stream);
int dtwa_destroy(dtwa_ctx *ctx);
/* name[256]; sm_count, clock_khz, cc (major*10+minor) */
int dtwa_device_info(dtwa_ctx *ctx, int which, char *name, int *sm_count, int *clock_khz, int *cc);
/* tuning: threads per CTA (0 = default) and resident CTAs per SM (0 = occupancy query) */
int dtwa_set_launch(dtwa_ctx *ctx, int block_threads, int ctas_per_sm);

/* The observables are compiled the way the reference compiles them (it evals a Julia lambda per
 * observable, src/TWA.jl:298): mode 1 specialises the ensemble kernel for the reducer set with NVRTC
 * (libnvrtc is dlopen'ed; ~1 s per new reducer set, cached per process), mode 0 interprets the
 * reducer program inside the ahead-of-time kernel, mode 2 (default) specialises when N*nt >= 2e9.
 * Both paths execute the same correctly rounded operations in the same order: identical results. */
int dtwa_set_jit(dtwa_ctx *ctx, int mode);
/* Small calls are specialised by a background host thread (see dtwa_set_jit).  Call this before the process exits
 * (atexit): it waits for compilations in flight and starts no new ones, so that no thread is inside libnvrtc while the
 * process tears it down. */
int dtwa_jit_quiesce(void);

/* compile-only check of the specialisation for `nexpr` averaged observables (no GPU needed) */
int dtwa_jit_compile(const dtwa_system *sys, const dtwa_solver *sol, const char *const *exprs, int32_t nexpr, char *src_out,
                     int64_t src_cap, int64_t *cubin_bytes);

/* ---- a5: ClassicalDicke.hamiltonian / ClassicalLMG.hamiltonian (src/ClassicalDicke.jl:80-88,
 *      src/ClassicalLMG.jl:63-70) on a batch u[n][nvar] -> h[n] ---- */
int dtwa_hamiltonian(dtwa_ctx *ctx, const dtwa_system *sys, const double *u, int64_t n, double *h);

/* ---- a7: ClassicalSystems.integrate (src/ClassicalSystems.jl:78-140) for n independent initial
 *      conditions, states reported at ts[nt] (sorted, >= 0, t0 = 0):
 *      out[n][nt][nvar] (NaN after a failure), status[n] (DTWA_TRAJ_*), nsteps[n] (accepted +
 *      rejected attempts; may be NULL). ---- */
int dtwa_integrate(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n,
                   const double *ts, int32_t nt, double *out, int32_t *status, int64_t *nsteps);

/* ---- integrate(...; get_fundamental_matrix=true, fundamental_matrix_initial=fm0)
 *      (src/ClassicalSystems.jl:110-128): the variational system d/dt (x, Phi) = (f(x), J(x) Phi) for n
 *      independent initial conditions; abstol=reltol=tol act on all nvar + nvar^2 components, like the
 *      reference's ArrayPartition.  fm0[n][nvar*nvar] column-major (Julia layout) or NULL = identity.
 *      out_u[n][nt][nvar], out_fm[n][nt][nvar*nvar] column-major. ---- */
int dtwa_integrate_fm(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, const double *fm0,
                      int64_t n, const double *ts, int32_t nt, double *out_u, double *out_fm, int32_t *status,
                      int64_t *nsteps);

/* ---- ClassicalSystems.lyapunov_spectrum (src/ClassicalSystems.jl:183-224) for n
 *      initial conditions at once (a Lyapunov map, docs/src/ClassicalDickeExamples.md:133-236): integrate
 *      (x, Phi = U) over intervals of length t_interval, modified Gram-Schmidt, lambda_i = sum log|v_i| / (k t),
 *      stop when |lambda_old - lambda| < lam_reltol |lambda| + lam_tol (tested after every column, as the
 *      reference does).  lambda[n][nvar]; x_final[n][nvar] (may be NULL); iters[n]; status[n] is
 *      DTWA_TRAJ_* or DTWA_LYAP_MAXITERS; nsteps[n] (may be NULL); kernel_ms (may be NULL). ---- */
int dtwa_lyapunov(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n,
                  double t_interval, double lam_tol, double lam_reltol, int32_t maxiters, double *lambda, double *x_final,
                  int32_t *iters, int32_t *status, int64_t *nsteps, double *kernel_ms);

/* ---- integrate(...; callback=ContinuousCallback(condition, affect; save_positions=(false,false)))
 *      as the docs use it for Poincare sections (docs/src/ClassicalDickeExamples.md:82-109,180-184): n initial
 *      conditions are integrated from 0 to t_end and every crossing of coef . u - offset = 0 is recorded
 *      (direction 0: both, +1: increasing, -1: decreasing).  The crossing is located to integrator accuracy
 *      (Newton on the crossing time with the integrator's own step from the start of the crossing step, then
 *      one classical RK4 Henon step -- the event function as independent variable -- so that g = 0 to rounding); the
 *      reference searches a root of its interpolant.  ev_t[n][max_events], ev_u[n][max_events][nvar], ev_dir[n][max_events] (+1/-1),
 *      ev_count[n] (keeps counting past max_events; entries of row i at or beyond min(ev_count[i], max_events)
 *      are unspecified), u_end[n][nvar], status[n], nsteps[n] (may be NULL). ---- */
typedef struct dtwa_event {
    double coef[4];
    double offset;
    int32_t direction;
    int32_t pad;
} dtwa_event;
int dtwa_integrate_events(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n,
                          double t_end, const dtwa_event *ev, int32_t max_events, double *ev_t, double *ev_u, int32_t *ev_dir,
                          int32_t *ev_count, double *u_end, int32_t *status, int64_t *nsteps, double *kernel_ms);

/* ---- the same with packed (CSR) output, for large batches: row i owns the entries [offsets[i], offsets[i+1]) of the
 *      event arrays, offsets[i+1] - offsets[i] = min(ev_count[i], max_events); offsets[n+1].  Two calls, because the
 *      caller can size its buffers only when the total offsets[n] is known: dtwa_integrate_events_packed solves and
 *      compacts on the device, dtwa_integrate_events_packed_fetch(total = offsets[n]) copies ev_t[total],
 *      ev_u[total][nvar], ev_dir[total].  The packed arrays stay valid until the next call on this context that
 *      integrates adaptive ensembles or events.  (No counterpart in the reference, which collects events in Julia
 *      arrays inside `affect!`.) ---- */
int dtwa_integrate_events_packed(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const double *u0, int64_t n,
                                 double t_end, const dtwa_event *ev, int32_t max_events, int64_t *offsets, int32_t *ev_count,
                                 double *u_end, int32_t *status, int64_t *nsteps, double *kernel_ms);
int dtwa_integrate_events_packed_fetch(dtwa_ctx *ctx, int64_t total, double *ev_t, double *ev_u, int32_t *ev_dir);

/* ---- EnergyShellProjections.∫∫dqdpδϵ (src/EnergyShellProjections.jl:49-96) for ncell cells
 *      (Q, P) at once (matrix_QP∫∫dqdpδϵ :159-199, energy_shell_average :476-508) and classical functions f
 *      given as 1..8 expressions over Q,q,P,p: out[ncell][nexpr] = sum over the roots q_+- and the n =
 *      2 floor(p_+/p_res) + 1 Chebyshev-Gauss nodes of (pi/(n w)) f([Q, q, P, p]); NaN and valid = 0 where the
 *      reference returns `nonvalue`.  onlyqroot: 0 both, +1, -1.  nodes[ncell] (may be NULL). ---- */
int dtwa_shell_quadrature(dtwa_ctx *ctx, const dtwa_system *sys, double eps, const double *QP, int64_t ncell,
                          const char *const *exprs, int32_t nexpr, double p_res, int32_t onlyqroot, double *out,
                          int32_t *valid, int32_t *nodes, double *kernel_ms);

/* ---- a8-a10: distribution.sample() / distribution.probability_density (src/TWA.jl:692-698,
 *      719-743,757-762) on the device: u[n][nvar] for trajectories offset..offset+n-1 ---- */
int dtwa_sample(dtwa_ctx *ctx, const dtwa_sampler *smp, int64_t n, uint32_t attempt, double *u);
int dtwa_pdf(dtwa_ctx *ctx, const dtwa_sampler *smp, const double *u, int64_t n, double *w);

/* ---- a12/a17 test hooks: evaluate an observable expression and scale_to_index (src/TWA.jl:53-64,
 *      277-299) on the device.  idx is 1-based, 0 = dropped. ---- */
int dtwa_eval_expr(dtwa_ctx *ctx, const dtwa_system *sys, const char *expr, const double *u, int64_t n, double *v);
int dtwa_scale_to_index(dtwa_ctx *ctx, const double *v, int64_t n, const dtwa_range *r, int64_t *idx);

/* ---- a13-a20: TWA.monte_carlo_integrate (src/TWA.jl:100-224) with the reducers of average
 *      (:565-568), variance (:630-633, composed by the shim), survival_probability (:677-680) and
 *      calculate_distribution (:552-556); nred > 1 is mcs_chain.
 *      tolerate_errors: failed trajectories are resampled (attempt+1) in follow-up rounds and the
 *      replacement contributes from the failure index on (src/TWA.jl:186-209). ---- */
int dtwa_ensemble(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const dtwa_sampler *smp, int64_t N,
                  const double *ts, int32_t nt, const dtwa_reducer *red, int32_t nred, int32_t tolerate_errors,
                  dtwa_result *res);

/* Split form for one-process-per-GPU jobs: launch leaves the reduced accumulators of this rank in
 * two device arenas (f64 sums, u64 counts+counters) and returns when they are complete; the host
 * framework all-reduces them (NCCL ncclSum) and makes the result visible (stream/event sync) before
 * fetch unpacks them into the caller's buffers. */
int dtwa_ensemble_launch(dtwa_ctx *ctx, const dtwa_system *sys, const dtwa_solver *sol, const dtwa_sampler *smp,
                         int64_t N, const double *ts, int32_t nt, const dtwa_reducer *red, int32_t nred,
                         int32_t tolerate_errors, dtwa_plan **plan);
int dtwa_plan_arenas(dtwa_plan *plan, void **d_f64, int64_t *n_f64, void **d_u64, int64_t *n_u64);
int dtwa_plan_fetch(dtwa_plan *plan, dtwa_result *res);
int dtwa_plan_free(dtwa_plan *plan);

/* ---- K0: FP64 roofline denominator: a DFMA-chain micro-benchmark on every SM; returns the
 *      achieved TFLOP/s (FMA = 2 flops) and the kernel time ---- */
int dtwa_fp64_peak(dtwa_ctx *ctx, double *tflops, double *ms);
/* K0b: the same with three distinct register operands per DFMA (the practical ceiling of an ODE kernel) */
int dtwa_fp64_peak_regs(dtwa_ctx *ctx, double *tflops, double *ms);
/* K0c: the DFMA rate that `chains` independent chains per thread reach with `warps_per_sm` one-warp CTAs per SM --
 * the shape of the adaptive kernels (3 warps per scheduler, 4 chains) without anything else; nreg = distinct vector-register
 * operands per DFMA (1: x*a+b with uniform a, b; 2: x*a+y, the form of a stage sum; 3: x*z+y).  A profiling aid. */
int dtwa_fp64_chains(dtwa_ctx *ctx, int chains, int warps_per_sm, int nreg, double *tflops, double *ms_out);
/* K0d: K0's DFMA chains with `alu_per_8` (0, 2, 4, 8, 16) instructions of another pipe woven in per 8 DFMAs (kind 0: integer
 * add, 1: FP32 FMA, 2: MUFU) -- whether the instructions around the FP64 arithmetic of a kernel cost DFMA issue slots.
 * Reports the DFMA rate alone.  A profiling aid. */
int dtwa_fp64_mix(dtwa_ctx *ctx, int alu_per_8, int kind, int warps_per_sm, double *tflops, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* DICKETWA_H */
