// dtwa_kernels.cuh -- the sm_100a kernels of libdicketwa.
//
//   K1/K2 ensemble_kernel<SYS,ALG>   sample -> integrate -> observables -> per-time reduction, fused
//                                    (replaces src/TWA.jl:163-213 + src/ClassicalSystems.jl:135-138 +
//                                    the reducers src/TWA.jl:336-370,413-491,652-668)
//   K3    integrate_kernel<SYS,ALG>  many independent trajectories, dense [n][nt][nvar] output
//                                    (src/ClassicalSystems.jl:78-140)
//   K4    the ts == [0.0] path is the same ensemble kernel with an all-zero step table
//                                    (src/TWA.jl:177,187-188)
//   K0    fp64_peak_kernel           DFMA-chain micro-benchmark, the roofline denominator
//   aux   finalize_sums_kernel (fixed-order merge of per-CTA partial sums), sample/pdf/eval/bin hooks
#pragma once
#include <type_traits>
#include "dtwa_device.cuh"

namespace dtwa {

constexpr int ST_OK = DTWA_TRAJ_OK, ST_DOMAIN = DTWA_TRAJ_DOMAIN, ST_MAXITERS = DTWA_TRAJ_MAXITERS,
              ST_BAD_IC = DTWA_TRAJ_BAD_IC, ST_INACTIVE = 4;

struct SolverArgs {
    SysPar par;
    // Uniform output grids (Julia ranges -- every docs example): all intervals share ONE row, which then
    // lives in kernel-parameter (constant-bank) space.  That matters: a DFMA with three distinct
    // vector-register operands issues every 3 cycles (register-bank reads), one with a constant-bank
    // operand every 2 (dtwa_fp64_peak_regs 24.7 vs dtwa_fp64_peak 34.2 TFLOP/s), and 28 of the 60
    // DFMAs of an RK4 step multiply by a step constant.  Per-interval rows loaded from `tab` end up in
    // vector registers (ptxas does not keep loaded values in uniform registers).
    StepTab uni;
    int32_t uniform;     // 1: intervals k >= 1 all use `uni`
    int32_t first_n;     // sub-steps of interval 0 in uniform mode (0 when ts[0] == 0)
    const StepTab *tab;  // [nt] fixed-step table (general case)
    const double *ts;    // [nt]
    double tol, dtmin, tend;
    long long maxiters;
    int32_t nt;
    int32_t interp;      // adaptive pairs: 1 = outputs from the Hermite interpolant over un-clipped steps (DTWA_OUTPUT_HERMITE)
    CtrlPar ctrl;        // step-size controller of the adaptive pairs
};
constexpr int INTERP_MAX_OUT = 32;   // an interpolating step is clipped so that it spans at most this many output times

template <int ALG>
struct is_fixed {
    static constexpr bool value = (ALG == DTWA_ALG_RK4 || ALG == DTWA_ALG_RK8);
};
struct Empty {};

// One trajectory in registers.
template <int SYS, int ALG>
struct Traj {
    static constexpr int NV = Sys<SYS>::NV;
    static constexpr bool FIXED = is_fixed<ALG>::value;
    double u[NV];
    int status;
    uint32_t nacc, nrej;
    typename std::conditional<FIXED, Empty, Adaptive<SYS, ALG>>::type ad;

    __device__ __forceinline__ void begin(const SolverArgs &S)
    {
        nacc = 0;
        nrej = 0;
        if (status == ST_OK && state_bad<SYS>(u)) status = ST_BAD_IC;
        if constexpr (!FIXED) {
            if (status == ST_OK) ad.init(S.par, u, S.tend, S.tol, S.ctrl);
        }
    }

    // integrate from the previous output time to ts[k]; returns the number of step iterations
    // this lane executed (for the lane-efficiency counter)
    __device__ __forceinline__ uint32_t advance(const SolverArgs &S, int k)
    {
        if constexpr (FIXED) {
            int n;
            if (S.uniform) {
                n = (k == 0) ? S.first_n : S.uni.n;
#pragma unroll 1
                for (int s = 0; s < n; ++s) {
                    if constexpr (ALG == DTWA_ALG_RK4)
                        rk4_step<SYS>(S.par, u, S.uni);  // constants are constant-bank operands
                    else
                        rk8_step<SYS>(S.par, u, S.uni.h);
                }
            } else {
                const StepTab r = S.tab[k];
                n = r.n;
#pragma unroll 1
                for (int s = 0; s < n; ++s) {
                    if constexpr (ALG == DTWA_ALG_RK4)
                        rk4_step<SYS>(S.par, u, r);
                    else
                        rk8_step<SYS>(S.par, u, r.h);
                }
            }
            if (status == ST_OK) nacc += (uint32_t)n;
            return (uint32_t)n;
        } else {
            uint32_t iters = 0;
            if (status == ST_OK) {
                const double tstop = S.ts[k];
#pragma unroll 1
                while (ad.t < tstop) {
                    if ((long long)(nacc + nrej) >= S.maxiters) {
                        status = ST_MAXITERS;
                        break;
                    }
                    bool accepted;
                    const int f = ad.attempt(S.par, u, tstop, S.tol, S.dtmin, accepted, S.ctrl);
                    ++iters;
                    if (f) {
                        status = ST_DOMAIN;
                        break;
                    }
                    if (accepted)
                        ++nacc;
                    else
                        ++nrej;
                }
            }
            return iters;
        }
    }
};

// ------------------------------------------------------------------------------------------------
struct FailRec {
    uint64_t idx;  // trajectory index relative to the sampler offset
    int32_t k;     // output index at which it failed
    int32_t pad;
};

struct HistDesc {
    int32_t two_d;       // 0: one expression axis (index register 0) against time; 1: two expression axes
    int32_t staged_off;  // offset (u32 elements) of the per-CTA shared-memory staging bins (1-D: per window slot; 2-D, all
                         // times in one matrix: one tile for the CTA's whole life), -1 = straight to global
    int64_t base;        // element offset in the u64 arena
    int64_t tstride;     // elements per output time (0: all times share one matrix)
    int32_t n0, n1;      // bins along index register 0 (x) / 1 (y)
};

// counters at the head of the u64 arena
enum { CNT_ACC = 0, CNT_REJ, CNT_LANE, CNT_FAILED, CNT_UNFINISHED, CNT_RESERVED5, CNT_RESERVED6, CNT_RESERVED7, CNT_N };

struct EnsArgs {
    SolverArgs S;
    SamplerPar smp;
    long long N;       // trajectories in this launch
    uint32_t attempt;  // resampling round (0 = first draw)
    int32_t nslots, nins, ncst, nrange, nhist, stage_bins;
    int32_t window;    // output times staged in shared memory between two CTA-wide flushes
    int32_t stage2d_bins;  // u32 cells of the per-CTA tiles of small 2-D histograms (fixed-step kernels)
    int32_t vgrid;         // adaptive kernels: number of warp STREAMS of the launch (= CTAs, or 2 x CTAs minus at most one for the pair kernels)
    const FailRec *redo;  // NULL in the main pass; else trajectory j re-draws redo[j].idx and contributes from redo[j].k
    FailRec *fail_out;
    unsigned *fail_count;
    unsigned fail_cap;
    const VmIns *prog;
    const double *consts;
    const RangeD *ranges;
    const HistDesc *hists;
    double *ring;             // adaptive kernels: [grid][window][32 lanes][NVO] ring of raw states (global memory)
    double *partial;          // [grid][nt][nslots]
    unsigned long long *u64;  // [CNT_N | n_valid[nt] | histograms...]
};

// Per-CTA shared-memory view used by the reducer VM.  Staging is per WINDOW of output times:
//   wsum  [rows][window][nslots]    observable values: one row per warp (fixed step: shuffle-reduced
//                                   first) or one row per thread (adaptive: lanes sit at different
//                                   output times, see ensemble_body), plain stores
//   stage [window][stage_bins]      1-D histogram bins (shared-memory integer atomics)
//   nvalid[window]                  contributing trajectories per time
// A CTA-wide barrier + flush happens once per window, not once per output time, so the warps of a
// CTA drift by up to `window` outputs and overlap each other's reduction latency.
struct VmShared {
    const double *consts;
    const RangeD *ranges;
    const HistDesc *hists;
    const VmIns *prog;
    double *wsum;
    unsigned *stage;
    unsigned *nvalid;
    unsigned *stage2d;
    int nslots, stage_bins, nwarps, window;
};

__host__ __device__ inline size_t ens_smem_bytes(int ncst, int nrange, int nhist, int nslots, int stage_bins, int nins,
                                                 int rows, int window, int stage2d_bins = 0)
{
    size_t b = 0;
    b += sizeof(double) * (size_t)ncst;
    b += sizeof(RangeD) * (size_t)nrange;
    b += sizeof(HistDesc) * (size_t)nhist;
    b += sizeof(double) * (size_t)rows * (size_t)window * (size_t)(nslots > 0 ? nslots : 1);
    b += sizeof(unsigned) * (size_t)window * (size_t)stage_bins;
    b += sizeof(unsigned) * (size_t)window;
    b += sizeof(unsigned) * (size_t)stage2d_bins;
    b += sizeof(VmIns) * (size_t)nins;
    return (b + 15) & ~size_t(15);
}

// The transcendental operations of the VM, behind a call: inlined into the interpreter's switch they spread its
// cases over tens of KB of code, and a small latency-bound ensemble then waits for the instruction cache at
// almost every VM instruction.  (Same libdevice routines as the inlined / specialised forms: same bits.)
__device__ __noinline__ double vm_slow_math(int op, double t, double acc)
{
    switch (op) {
    case OP_POW: return pow(t, acc);
    case OP_ATAN2: return atan2(t, acc);
    case OP_SIN: return sin(acc);
    case OP_COS: return cos(acc);
    case OP_TAN: return tan(acc);
    case OP_EXP: return exp(acc);
    case OP_LOG: return log(acc);
    case OP_ACOS: return acos(acc);
    case OP_ASIN: return asin(acc);
    case OP_ATAN: return atan(acc);
    default: return acc;
    }
}

// Evaluate the reducer program for one thread at output index k (slot w of the current window).
// Control flow is uniform across the CTA (the program is shared); `valid` masks this thread's
// contributions.  Accumulator + 4 stack entries live in registers.
template <int NV, bool PERLANE>
__device__ __forceinline__ void vm_run(const VmShared &V, const SamplerPar *smp, unsigned long long *u64, const double x0,
                                       const double x1, const double x2, const double x3, const bool valid, const int k,
                                       const int w, double *store_out, const bool emit = true)
{
    // emit = false: this lane only keeps the warp company (the program is run by the whole warp so that its per-instruction
    // loop never splits the warp); it stores and counts nothing.  valid implies emit.
    double acc = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    double ov[VM_STACK - VM_REG_STACK];
    int depth = 0;
    int ireg0 = 0, ireg1 = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#define VM_PUSH()                               \
    do {                                        \
        if (depth >= VM_REG_STACK) ov[depth - VM_REG_STACK] = s3; \
        s3 = s2; s2 = s1; s1 = s0; s0 = acc;    \
        ++depth;                                \
    } while (0)
#define VM_POP(dst)                             \
    do {                                        \
        dst = s0; s0 = s1; s1 = s2; s2 = s3;    \
        --depth;                                \
        if (depth >= VM_REG_STACK) s3 = ov[depth - VM_REG_STACK]; \
    } while (0)
#define VM_VAR(a) (((a) == 0) ? x0 : ((a) == 1) ? x1 : ((a) == 2) ? x2 : x3)
    VmIns in = V.prog[0];
#pragma unroll 1
    for (int pc = 1; in.op != OP_END; ++pc) {
        const VmIns nxt = V.prog[pc];  // prefetch: the program always ends with OP_END
        double t;
        switch (in.op) {
        case OP_PLDV: VM_PUSH();
        case OP_LDV: acc = VM_VAR(in.a); break;
        case OP_PLDC: VM_PUSH();
        case OP_LDC: acc = V.consts[in.b]; break;
        case OP_PLDPDF: VM_PUSH();
        case OP_LDPDF: {
            double uu[NV];
            uu[0] = x0;
            uu[1] = x1;
            if constexpr (NV == 4) {
                uu[2] = x2;
                uu[3] = x3;
            }
            acc = pdf_eval<NV>(*smp, uu);
            break;
        }
        case OP_ADD: VM_POP(t); acc = __dadd_rn(t, acc); break;
        case OP_SUB: VM_POP(t); acc = __dsub_rn(t, acc); break;
        case OP_MUL: VM_POP(t); acc = __dmul_rn(t, acc); break;
        case OP_DIV: VM_POP(t); acc = __ddiv_rn(t, acc); break;
        case OP_POW:
        case OP_ATAN2: VM_POP(t); acc = vm_slow_math(in.op, t, acc); break;
        case OP_MIN: VM_POP(t); acc = fmin(t, acc); break;
        case OP_MAX: VM_POP(t); acc = fmax(t, acc); break;
        case OP_ADDC: acc = __dadd_rn(acc, V.consts[in.b]); break;
        case OP_SUBC: acc = __dsub_rn(acc, V.consts[in.b]); break;
        case OP_RSUBC: acc = __dsub_rn(V.consts[in.b], acc); break;
        case OP_MULC: acc = __dmul_rn(acc, V.consts[in.b]); break;
        case OP_DIVC: acc = __ddiv_rn(acc, V.consts[in.b]); break;
        case OP_RDIVC: acc = __ddiv_rn(V.consts[in.b], acc); break;
        case OP_ADDV: acc = __dadd_rn(acc, VM_VAR(in.a)); break;
        case OP_SUBV: acc = __dsub_rn(acc, VM_VAR(in.a)); break;
        case OP_RSUBV: acc = __dsub_rn(VM_VAR(in.a), acc); break;
        case OP_MULV: acc = __dmul_rn(acc, VM_VAR(in.a)); break;
        case OP_DIVV: acc = __ddiv_rn(acc, VM_VAR(in.a)); break;
        case OP_RDIVV: acc = __ddiv_rn(VM_VAR(in.a), acc); break;
        case OP_NEG: acc = -acc; break;
        case OP_SQR: acc = __dmul_rn(acc, acc); break;
        case OP_CUBE: acc = __dmul_rn(__dmul_rn(acc, acc), acc); break;
        case OP_IPOW: acc = ipow_rn(acc, (int)(int16_t)in.b); break;
        case OP_SQRT: acc = __dsqrt_rn(acc); break;
        case OP_SIN:
        case OP_COS:
        case OP_TAN:
        case OP_EXP:
        case OP_LOG:
        case OP_ACOS:
        case OP_ASIN:
        case OP_ATAN: acc = vm_slow_math(in.op, 0.0, acc); break;
        case OP_ABS: acc = fabs(acc); break;
        case OP_SUM: {
            double v = valid ? acc : 0.0;
            if constexpr (PERLANE) {   // adaptive kernels' flush: every lane accumulates one output time in its own row
                if (emit) V.wsum[(size_t)threadIdx.x * V.nslots + in.b] += v;
            } else {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
                if (lane == 0) V.wsum[((size_t)warp * V.window + w) * V.nslots + in.b] = v;
            }
            break;
        }
        case OP_BIN: {
            const int idx = scale_to_index(acc, V.ranges[in.b]);
            if (in.a == 0)
                ireg0 = idx;
            else
                ireg1 = idx;
            break;
        }
        case OP_HIST: {
            const HistDesc h = V.hists[in.b];
            if (!h.two_d) {
                if (valid && ireg0 > 0) {
                    if (h.staged_off >= 0)
                        atomicAdd(&V.stage[(size_t)w * V.stage_bins + h.staged_off + (int)(ireg0 - 1)], 1u);
                    else
                        atomicAdd(&u64[h.base + (long long)k * h.tstride + (ireg0 - 1)], 1ull);
                }
            } else {
                if (valid && ireg0 > 0 && ireg1 > 0) {
                    const unsigned cell = (unsigned)(ireg1 - 1) * (unsigned)h.n0 + (unsigned)(ireg0 - 1);   // n0 * n1 < 2^31 (checked)
                    if (h.staged_off >= 0)
                        atomicAdd(&V.stage2d[h.staged_off + cell], 1u);
                    else
                        atomicAdd(&u64[h.base + (long long)k * h.tstride + cell], 1ull);
                }
            }
            break;
        }
        case OP_STORE:
            if (store_out) *store_out = acc;
            break;
        default: break;
        }
        in = nxt;
    }
#undef VM_PUSH
#undef VM_POP
#undef VM_VAR
}

#ifndef DTWA_RK8_MAX_THREADS
#define DTWA_RK8_MAX_THREADS 128
#endif
#if defined(DTWA_ADAPT_PAIR) && defined(DTWA_JIT_UNIT)
#define DTWA_PAIR_UNIT 1   // this translation unit is a run-time specialisation with two warp streams per warp
#endif
#ifndef DTWA_ADAPT_MIN_BLOCKS
#ifdef DTWA_PAIR_UNIT
#define DTWA_ADAPT_MIN_BLOCKS 8    // two warps per scheduler, <= 255 registers: two trajectories per lane
#else
#define DTWA_ADAPT_MIN_BLOCKS 12
#endif
#endif
#ifndef DTWA_DP5_MIN_BLOCKS
#ifdef DTWA_JIT_UNIT
#define DTWA_DP5_MIN_BLOCKS (DTWA_ADAPT_MIN_BLOCKS == 12 ? 16 : DTWA_ADAPT_MIN_BLOCKS)   // the 7-stage pair fits 128 registers without spills: 4 warps per scheduler (+3.6 %)
#else
#define DTWA_DP5_MIN_BLOCKS DTWA_ADAPT_MIN_BLOCKS   // (the interpreter build needs the registers for the reducer VM)
#endif
#endif
// Register budgets.  RK4 fits 64 registers/thread: 4 CTAs of 256 threads per SM.  The embedded pairs run ONE warp per CTA
// (adaptive_stream); the register file is split over the four SM sub-partitions (16384 registers each), so the only
// occupancies there are 3 warps per sub-partition at <= 168 registers (12 CTAs/SM) or 4 at <= 128 (16 CTAs/SM, with spills
// for the 12-stage pair).  Without a bound ptxas takes 170-213 registers and lands on 2 warps per sub-partition.
// DTWA_ADAPT_WARPS: warp streams per CTA of the adaptive kernels.  Every warp is still its own "virtual CTA" (stream number
// vcta = blockIdx.x * DTWA_ADAPT_WARPS + warp: its own ring, its own partial row, its own trajectories -- the same ones as
// CTA vcta of the one-warp launch, so results are bit-identical); what changes is how the hardware places them: the warps of
// one CTA go to the four schedulers of an SM in turn, one-warp CTAs go wherever a warp slot is free.
#ifndef DTWA_ADAPT_WARPS
#define DTWA_ADAPT_WARPS 1
#endif
template <int ALG>
struct ens_launch {
    static constexpr int adapt_blocks = (ALG == DTWA_ALG_DP5 || ALG == DTWA_ALG_TSIT5) ? DTWA_DP5_MIN_BLOCKS : DTWA_ADAPT_MIN_BLOCKS;
    static constexpr int max_threads = (ALG == DTWA_ALG_RK4) ? 256 : (ALG == DTWA_ALG_RK8) ? DTWA_RK8_MAX_THREADS : 32 * DTWA_ADAPT_WARPS;
    static constexpr int min_blocks = (ALG == DTWA_ALG_RK4) ? 4 : (ALG == DTWA_ALG_RK8) ? 1 : (adapt_blocks + DTWA_ADAPT_WARPS - 1) / DTWA_ADAPT_WARPS;
};

// Reducers compiled at run time for one reducer set (dtwa_jit.hpp generates the definition: the same
// operation sequence as vm_run, as straight-line code).  Declared here, defined only in JIT units.
template <int NV, bool PERLANE>
__device__ __forceinline__ void jit_reduce(const VmShared &V, const SamplerPar *smp, unsigned long long *u64, const double x0,
                                           const double x1, const double x2, const double x3, const bool valid, const int k,
                                           const int w, const bool emit = true);   // emit = false: compute, store nothing (valid implies emit)

// Adaptive pairs: ONE warp per CTA (the host launches 32 threads).  Each lane integrates its own STREAM of
// trajectories -- lane l of CTA c takes trajectories (c*32 + l) + m*gridDim*32, m = 0, 1, ... -- and numbers the output
// times of the stream g = m*nt + k.  In every iteration of the flat loop each lane that owes a step attempts ONE
// step; a lane that has reached its next output time stores its STATE (the observables' coordinates, 8 NVO bytes) into
// its cell of a ring of R <= 256 outputs in global memory (A.ring, layout [w][lane][NVO]; 0.5 GB at most -- larger than
// L2, it streams through DRAM at ~60 GB/s, 1 % of the HBM bandwidth) and, at the end of a trajectory, draws the next one
// without waiting for the other lanes.  Observables and reducers do NOT run when a lane emits -- lanes emit at different
// iterations, so that code would run in almost every iteration with 2-3 active lanes (ncu: 93 % of the iterations, 8-12 %
// of a warp's time) -- but when the ring is flushed: in halves, in the order of g, as soon as the slowest lane has
// passed a half, the whole warp reads one output time per step (a coalesced 1 KB row), every lane evaluates the reducer
// program for its cell, sums go through the shuffle tree of the fixed-step kernels and one RED per (time, slot) into the
// CTA's partial row -- the same values in the same order in every run.  An empty cell (lane failed, skipped, ran out of
// trajectories, or re-draws a trajectory that contributes only from a later time on) holds a NaN marker.  A lane only
// waits when it is more than R/2 .. R outputs ahead of the slowest lane of its warp.
template <int SYS, int ALG, bool JIT>
__device__ __forceinline__ void adaptive_stream(const EnsArgs &A, const VmShared &V, double *my_partial,
                                                unsigned long long *nvalid_g, double *ring, const int vcta)
{
    constexpr int NV = Sys<SYS>::NV, NVO = Sys<SYS>::NVO;
    const int lane = threadIdx.x & 31;
    const int nt = A.S.nt, W = A.window;
    const int H = (W >= 2) ? W / 2 : 1, R = (W >= 2) ? 2 * H : 1;   // half and ring size
    const long long stride = (long long)A.vgrid * 32;                 // (vgrid warp streams: one per warp of every CTA)
    long long j = (long long)vcta * 32 + lane;                        // current trajectory of this lane
    long long g = 0, flushed = 0;                                     // stream positions: next output of this lane / flushed
    int k = nt;                                                       // next output index of the current trajectory (nt: none)
    int wslot = 0;                                                    // g % R, kept incrementally (no 64-bit division in the loop)
    int start_k = 0;
    uint64_t idx = 0;
    bool have = false;                                                // a trajectory is in flight
    double tstop = 0.0, tnext = 0.0;                                  // the next output time and the one after it (prefetched: a lane that
                                                                      // lands on ts[k] must not wait for the load of ts[k+1])
    // The loop control of the 7-stage pair is written lean: the output mode is a compile-time constant of a specialised unit (the
    // other loop body is not generated), the maxiters test is a select, ts[k+1] is prefetched.  For the 12-stage pair the same
    // source makes ptxas hoist 16 coefficients into vector registers (168 registers + spills, 32 R2UR per attempt: 838
    // instructions and 2 080 statically scheduled cycles per attempt instead of 818 / 1 925), so it keeps the plain form.
#ifndef DTWA_LEAN_DOP853
#define DTWA_LEAN_DOP853 0   // (tuning experiments)
#endif
    constexpr bool LEAN = (ALG == DTWA_ALG_DP5 || ALG == DTWA_ALG_TSIT5) || (DTWA_LEAN_DOP853 != 0);
#ifdef DTWA_JIT_INTERP
    const bool interp_mode = LEAN ? (DTWA_JIT_INTERP != 0) : (A.S.interp != 0);
#else
    const bool interp_mode = A.S.interp != 0;
#endif
    uint32_t iters = 0;
    int budget = 0;                                                   // attempts left before maxiters (current trajectory)
    int tlim_k = -1;                                                  // Hermite outputs: the output index `tlim` was loaded for
    double tlim = 0.0;                                                // ... and the time a step is clipped at
    unsigned long long acc_steps = 0, rej_steps = 0;
    Traj<SYS, ALG> T;
    T.status = ST_INACTIVE;
    T.nacc = T.nrej = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) T.u[i] = 0.0;

    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    double *mycell = ring + (size_t)lane * NVO;     // this lane's cell of ring slot 0; slot w is 32 * NVO doubles further per w
    auto flush_half = [&](long long f0, long long f1) {   // stream outputs [f0, f1), executed by the whole warp, fixed order
        // lane l takes output time f0 + c + l of every chunk c of 32 and walks over the 32 lanes' cells of that time in lane
        // order: no shuffles, no barrier per output, a fixed summation order, one RED per (time, slot)
        __syncwarp();
        const int cnt = (int)(f1 - f0);
        const int w0 = (int)(f0 % R), k0 = (int)(f0 % nt);
        double *row = V.wsum + (size_t)threadIdx.x * A.nslots;
        for (int c = 0; c < cnt; c += 32) {
            const int o = c + lane;
            const bool mine = o < cnt;
            const int w = (w0 + o) % R, kq = (k0 + o) % nt;
            for (int sl = 0; sl < A.nslots; ++sl) row[sl] = 0.0;
            unsigned nv = 0;
            constexpr int RB = JIT ? 8 : 2;   // cells in flight per lane: the scattered 32-byte loads of a batch overlap (one DRAM round trip per
                                              // batch); the interpreter's reducer call is large, so its loop is unrolled only twice
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += RB) {
                double xa[RB][NVO];
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const double *cell = ring + ((size_t)w * 32 + (r0 + q)) * NVO;
                    if (mine) {
                        if constexpr (NVO == 4) {
                            const double2 a = __ldcg(reinterpret_cast<const double2 *>(cell)), b = __ldcg(reinterpret_cast<const double2 *>(cell + 2));
                            xa[q][0] = a.x; xa[q][1] = a.y; xa[q][2] = b.x; xa[q][3] = b.y;
                        } else {
                            const double2 a = __ldcg(reinterpret_cast<const double2 *>(cell));
                            xa[q][0] = a.x; xa[q][1] = a.y;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < NVO; ++i) xa[q][i] = qnan;
                    }
                }
#pragma unroll
                for (int q = 0; q < RB; ++q)
                    if (mine) ring[((size_t)w * 32 + (r0 + q)) * NVO] = qnan;   // the cells are free again
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const bool valid = (xa[q][0] == xa[q][0]);
                    nv += valid ? 1u : 0u;
                    if constexpr (JIT)
                        jit_reduce<NVO, true>(V, &A.smp, A.u64, xa[q][0], xa[q][1], NVO == 4 ? xa[q][NVO - 2] : 0.0, NVO == 4 ? xa[q][NVO - 1] : 0.0, valid, kq,
                                              0, valid);
                    else
                        vm_run<NVO, true>(V, &A.smp, A.u64, xa[q][0], xa[q][1], NVO == 4 ? xa[q][NVO - 2] : 0.0, NVO == 4 ? xa[q][NVO - 1] : 0.0, valid, kq, 0,
                                          nullptr, valid);
                }
            }
            if (mine) {
                for (int sl = 0; sl < A.nslots; ++sl) atomicAdd(&my_partial[(size_t)kq * A.nslots + sl], row[sl]);
                if (nv) atomicAdd(&nvalid_g[kq], (unsigned long long)nv);
            }
        }
        __syncwarp();
    };
    auto emit_state = [&](const double (&uo)[NV], bool valid, int wslot_) {   // the lane's state at an output time -> its ring cell
        if (!valid) return;
        double xo[NVO];
        from_chart<SYS>(uo, xo);
        double *cell = mycell + (size_t)wslot_ * (32 * NVO);
        if constexpr (NVO == 4) {
            *reinterpret_cast<double2 *>(cell) = make_double2(xo[0], xo[1]);
            *reinterpret_cast<double2 *>(cell + 2) = make_double2(xo[2], xo[3]);
        } else
            *reinterpret_cast<double2 *>(cell) = make_double2(xo[0], xo[1]);
    };

#pragma unroll 1
    while (true) {
        // ================= service part (cold): new trajectories, failures, flushes =================
        // a lane without a trajectory draws its next one (or retires)
        if (!have && j < A.N) {
            idx = (uint64_t)j;
            start_k = 0;
            if (A.redo) {
                const FailRec r = A.redo[j];
                idx = r.idx;
                start_k = r.k;
            }
            acc_steps += T.nacc;
            rej_steps += T.nrej;
            T.status = ST_OK;
            {
                double x[NVO];
                draw_initial<NVO>(A.smp, idx, A.attempt, x);
                to_chart<SYS>(x, T.u);
            }
            T.begin(A.S);
            have = true;
            k = 0;
            tstop = A.S.ts[0];
            if constexpr (LEAN) tnext = A.S.ts[nt > 1 ? 1 : 0];
            budget = (int)(A.S.maxiters > 0x7fffffffll ? 0x7fffffffll : A.S.maxiters);
            tlim_k = -1;
        }
        // failures end a trajectory: report, skip the rest of its outputs
        if (have && T.status != ST_OK) {
            bool queued = false;
            if (T.status != ST_MAXITERS) {
                atomicAdd(&A.u64[CNT_FAILED], 1ull);
                if (A.fail_out && A.smp.kind != DTWA_SAMPLER_HOST_ARRAY && A.smp.kind != DTWA_SAMPLER_DEVICE_ARRAY) {
                    const unsigned pos = atomicAdd(A.fail_count, 1u);
                    if (pos < A.fail_cap) {
                        FailRec r;
                        r.idx = idx;
                        r.k = k;
                        r.pad = 0;
                        A.fail_out[pos] = r;
                        queued = true;
                    }
                }
            }
            if (!queued) atomicAdd(&A.u64[CNT_UNFINISHED], 1ull);
            g += (long long)(nt - k);
            wslot = (int)(((long long)wslot + (nt - k)) % R);
            k = nt;
        }
        if (have && k >= nt) {   // trajectory finished (or abandoned): the next one is drawn in the following round
            have = false;
            j += stride;
        }
        // flush what every lane has passed; stop when nothing is left
        {
            const bool retired = !have && j >= A.N;
            const long long ahead = retired ? (long long)0x7fffffff : g - flushed;
            const int amin = __reduce_min_sync(0xffffffffu, (int)(ahead > 0x7fffffff ? 0x7fffffff : ahead));
            if (amin == 0x7fffffff) {   // every lane has retired: flush the tail of the stream
                const long long gend = __reduce_max_sync(0xffffffffu, (int)(g - flushed)) + flushed;
                while (flushed < gend) {
                    const long long f1 = (flushed + H < gend) ? flushed + H : gend;
                    flush_half(flushed, f1);
                    flushed = f1;
                }
                break;
            }
            if (amin >= H) {
                flush_half(flushed, flushed + H);
                flushed += H;
            }
        }
        if (__any_sync(0xffffffffu, !have && j < A.N)) continue;   // somebody still has to draw
        // ================= hot loop: one attempted step per lane and iteration =================
        // (lane-local 32-bit bookkeeping: `ahead` = outputs this lane is ahead of the flushed part of the ring, `budget` = attempts
        //  left before maxiters; one REDUX decides whether to leave)
        int ahead = have ? (int)(g - flushed) : 0;
#ifdef DTWA_PIPELINED_LOOP   // (opt-in experiment, measured 0-3 % SLOWER except DP5 +3.6 %: ptxas keeps the emit block behind a branch)
        if (JIT && !interp_mode) {
            // ---- software-pipelined form (specialised kernels, outputs by stepping onto ts[k]) ----
            // An attempt ends in a serial, latency-bound tail -- error norm -> controller -> commit -> emit -> vote -- during
            // which the warp feeds nothing to the FP64 pipe (30-40 % of its time with three warps per scheduler, ncu source
            // page of r2_ncu_dop853_cfg4_v6).  Only the accept decision and the next step size are needed by the next
            // attempt.  So the loop body is rotated: the output a lane landed on in iteration n-1 is emitted at the START of
            // iteration n, the exit vote taken there is acted on one iteration later, every lane runs the trial step
            // (inactive ones with h = 0) and the decisions are selects -- one straight-line block in which the scheduler
            // overlaps the emit, the vote and the bookkeeping with the stage arithmetic of the next attempt.  Same trial(),
            // same controller expressions as Adaptive::attempt: bit-identical trajectories.
            const CtrlPar &c = A.S.ctrl;
            bool pend = have && T.status == ST_OK && k < nt && ahead < R && !(T.ad.t < tstop);
            tnext = (have && k + 1 < nt) ? A.S.ts[k + 1] : tstop;
            unsigned vote = 2u;
#pragma unroll 1
            while (true) {
                if ((vote & 1u) || !(vote & 2u)) break;
                ++iters;
                {   // [E] the output reached in the previous iteration (state T.u is still the one that landed on ts[k])
                    emit_state(T.u, pend && (k >= start_k), wslot);
                    if (pend) {
                        ++k;
                        ++ahead;
                        wslot = (wslot + 1 == R) ? 0 : wslot + 1;
                        tstop = tnext;
                        if (k + 1 < nt) tnext = A.S.ts[k + 1];   // prefetched: needed one output time from now
                    }
                }
                {   // [V] exit vote, acted on at the top of the next iteration
                    const bool service = have && (T.status != ST_OK || k >= nt);
                    const bool behind = have && ahead < H;
                    vote = __reduce_or_sync(0xffffffffu, (service ? 1u : 0u) | (behind ? 2u : 0u));
                }
                // [S] trial step (every lane; inactive ones take h = 0)
                const bool want = have && T.status == ST_OK && k < nt && ahead < R && T.ad.t < tstop;
                const bool active = want && budget > 0;
                if (want && budget <= 0) T.status = ST_MAXITERS;
                budget -= active ? 1 : 0;
                const bool forced = T.ad.habs < A.S.dtmin;
                const double habs_eff = forced ? A.S.dtmin : T.ad.habs;
                const bool last = (T.ad.t + habs_eff >= tstop);
                const bool clipped = last && ((tstop - T.ad.t) < habs_eff);
                double h = last ? tstop - T.ad.t : habs_eff;
                if (!active) h = 0.0;
                double yn[NV], kn[NV], err;
                float le;
                T.ad.trial(A.S.par, T.u, h, A.S.tol, yn, kn, err, le);
                // [C] decisions as selects
                const bool bad = state_bad<SYS>(yn) || !(err == err) || isinf(err);
#ifndef DTWA_LE_FROM_FACTORS
                le = lg2_approx(fminf(fmaxf((float)err, 1e-30f), 1e30f));
#endif
                const float e_rej = ex2_approx(-c.b1 * le);
                const float e_acc = ex2_approx(-fmaf(c.b1, le, c.nb2 * T.ad.lqold));
                float fac_acc = fminf(c.fac_max, fmaxf(c.fac_min_acc, c.g * e_acc));
                if (T.ad.rejected) fac_acc = fminf(fac_acc, c.post_reject_cap);
                const float fac_rej = fmaxf(c.fac_min_rej, c.g * e_rej);
                const bool acc = active && !bad && (forced || err < c.acc_thr);
                const bool fail = active && bad && forced;
                double hn = h * (double)(acc ? fac_acc : fac_rej);
                if (bad) hn = h * c.dom_shrink;
                if (acc && clipped && hn < habs_eff) hn = habs_eff;
                if (acc && !clipped) T.ad.lqold = fmaxf(le, c.lqoldinit);
                if (active && !fail) T.ad.habs = hn;
                if (active) T.ad.rejected = !acc;
                if (acc) {
                    T.ad.t = last ? tstop : T.ad.t + h;
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        T.u[i] = yn[i];
                        T.ad.k0[i] = kn[i];
                    }
                }
                if (fail) T.status = ST_DOMAIN;
                T.nacc += acc ? 1u : 0u;
                T.nrej += (active && !acc && !fail) ? 1u : 0u;
                pend = have && T.status == ST_OK && k < nt && ahead < R && !(T.ad.t < tstop);
            }
        } else
#endif
#pragma unroll 1
        while (true) {
            // leave when a lane needs service or a half of the ring can be flushed
            const bool service = have && (T.status != ST_OK || k >= nt);
            const bool behind = have && ahead < H;          // this lane still keeps the older half of the ring busy
#ifndef DTWA_VOTE_ANY
            const unsigned vote = __reduce_or_sync(0xffffffffu, (service ? 1u : 0u) | (behind ? 2u : 0u));
            if ((vote & 1u) || !(vote & 2u)) break;
#else   // (experiment, call 23: two VOTE.ANY straight into predicates -- 1 % slower than the one REDUX.OR)
            if (__any_sync(0xffffffffu, service) || !__any_sync(0xffffffffu, behind)) break;
#endif
            ++iters;
            if (!interp_mode) {
            const bool stepping = have && ahead < R;
            auto one_attempt = [&]() {
                --budget;
                bool accepted;
                if (T.ad.attempt(A.S.par, T.u, tstop, A.S.tol, A.S.dtmin, accepted, A.S.ctrl)) T.status = ST_DOMAIN;
                if (accepted)
                    ++T.nacc;
                else if (T.status == ST_OK)
                    ++T.nrej;
            };
            if constexpr (LEAN) {
                const bool want = stepping && T.ad.t < tstop;
                if (want && budget <= 0) T.status = ST_MAXITERS;
                if (want && budget > 0) one_attempt();
            } else {
                if (stepping && T.ad.t < tstop) {
                    if (budget <= 0)
                        T.status = ST_MAXITERS;
                    else
                        one_attempt();
                }
            }
            const bool reached = stepping && T.status == ST_OK && !(T.ad.t < tstop);   // this lane emits output k now
            if (reached) {
                emit_state(T.u, k >= start_k, wslot);
                ++k;
                ++ahead;
                wslot = (wslot + 1 == R) ? 0 : wslot + 1;
                if constexpr (LEAN) {
                    tstop = tnext;                                  // (== ts[k] whenever k < nt)
                    if (k + 1 < nt) tnext = A.S.ts[k + 1];
                } else if (k < nt)
                    tstop = A.S.ts[k];
            }
            } else {
            // ---- DTWA_OUTPUT_HERMITE: steps ignore the output times (clipped only at tlim = the INTERP_MAX_OUT-th output time
            // ahead, so that one step never fills more than INTERP_MAX_OUT ring slots, and at the last one); every output time
            // inside an accepted step is served from the Hermite interpolant over that step, right away ----
            const int maxout = min(INTERP_MAX_OUT, H);   // (H >= 32 whenever the ring holds >= 64 output times)
            const bool stepping = have && ahead <= R - maxout;
            bool accepted = false;
            double yprev[NV], fprev[NV];
            if (stepping && T.status == ST_OK && T.ad.t < tstop) {
                if (budget <= 0) {
                    T.status = ST_MAXITERS;
                } else {
                    --budget;
                    if (k != tlim_k) {   // (a global load: only when the next output index has moved)
                        tlim_k = k;
                        tlim = A.S.ts[min(k + maxout - 1, nt - 1)];
                    }
                    if (T.ad.template attempt<true>(A.S.par, T.u, tlim, A.S.tol, A.S.dtmin, accepted, A.S.ctrl, yprev, fprev)) T.status = ST_DOMAIN;
                    if (accepted)
                        ++T.nacc;
                    else if (T.status == ST_OK)
                        ++T.nrej;
                }
            }
            // outputs due: after an accepted step all ts[k] <= t (interpolated); before the first step the ones at t = 0
            bool due = stepping && T.status == ST_OK && k < nt && !(T.ad.t < tstop);
            while (due) {
                double uo[NV];
                if (accepted && tstop < T.ad.t) {
                    const double theta = (tstop - (T.ad.t - T.ad.hlast)) / T.ad.hlast;
                    hermite_interpolant<NV>(yprev, fprev, T.u, T.ad.k0, T.ad.hlast, theta, uo);
                } else {
#pragma unroll
                    for (int i = 0; i < NV; ++i) uo[i] = T.u[i];
                }
                if (state_bad<SYS>(uo)) {   // the interpolated state left the chart: the trajectory fails at this output
                    T.status = ST_DOMAIN;
                    break;
                }
                emit_state(uo, k >= start_k, wslot);
                ++k;
                ++ahead;
                wslot = (wslot + 1 == R) ? 0 : wslot + 1;
                if (k < nt) tstop = A.S.ts[k];
                due = k < nt && !(T.ad.t < tstop);
            }
            }
        }
        if (have) g = flushed + ahead;
    }
    // ---- step counters: one atomic per warp ----
    acc_steps += T.nacc;
    rej_steps += T.nrej;
    unsigned long long a = acc_steps, r = rej_steps;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        r += __shfl_xor_sync(0xffffffffu, r, off);
    }
    if (lane == 0) {
        atomicAdd(&A.u64[CNT_ACC], a);
        if (r) atomicAdd(&A.u64[CNT_REJ], r);
        atomicAdd(&A.u64[CNT_LANE], 32ull * iters);
    }
}

#ifdef DTWA_PAIR_UNIT
// adaptive_stream for TWO warp streams per physical warp (run-time specialised kernels, outputs by stepping onto ts[k]).
// The embedded pairs are latency-bound: their register budget allows three warps per scheduler, and one attempt is a chain of
// ~13 x 11 dependent FP64 instructions plus a serial tail (error norm -> controller -> commit -> emit -> vote); with three
// warps the FP64 pipe is 70 % busy (profiles/r2_ncu_dop853_cfg4.md).  Here every lane carries the trajectories of TWO streams
// -- warp stream 2 b + s of a launch with `vgrid` streams, s = 0, 1: the same trajectories, the same ring, the same partial
// row and the same summation order as CTA 2 b + s of the one-stream kernel launched on vgrid CTAs, hence bit-identical
// results -- and attempts one step of each per iteration through Adaptive::trial2, whose generated stage code interleaves the
// two dependency chains in one basic block.  Two resident warps per scheduler (<= 255 registers) then carry four streams
// with instruction-level parallelism 2 each.  Decisions are selects (the expressions of Adaptive::attempt).
template <int SYS, int ALG>
__device__ __forceinline__ void adaptive_stream_pair(const EnsArgs &A, const VmShared &V, unsigned long long *nvalid_g)
{
    constexpr int NV = Sys<SYS>::NV, NVO = Sys<SYS>::NVO;
    const int lane = threadIdx.x & 31;
    const int nt = A.S.nt, W = A.window;
    const int H = (W >= 2) ? W / 2 : 1, R = (W >= 2) ? 2 * H : 1;   // half and ring size
    const long long stride = (long long)A.vgrid * 32;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    const CtrlPar &c = A.S.ctrl;

    struct Stream {
        long long j, g, flushed;
        uint64_t idx;
        double tstop;
        double *ring, *partial;
        int k, wslot, start_k, budget;
        bool have, done;
        unsigned long long acc_steps, rej_steps;
    };
    Stream X[2];
    Traj<SYS, ALG> T[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const long long vc = 2ll * blockIdx.x + s;
        X[s].done = false;
        X[s].j = vc < A.vgrid ? vc * 32 + lane : A.N;   // (an odd number of streams: the last warp carries one)
        X[s].g = X[s].flushed = 0;
        X[s].idx = 0;
        X[s].tstop = 0.0;
        X[s].ring = A.ring + (size_t)vc * 32 * W * NVO;
        X[s].partial = A.partial + (size_t)vc * nt * A.nslots;
        X[s].k = nt;
        X[s].wslot = X[s].start_k = X[s].budget = 0;
        X[s].have = false;
        X[s].acc_steps = X[s].rej_steps = 0;
        T[s].status = ST_INACTIVE;
        T[s].nacc = T[s].nrej = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) T[s].u[i] = 0.0;
        T[s].ad.t = 0.0;
        T[s].ad.habs = 0.0;
        T[s].ad.lqold = c.lqoldinit;
        T[s].ad.rejected = false;
#pragma unroll
        for (int i = 0; i < NV; ++i) T[s].ad.k0[i] = 0.0;
        if (vc < A.vgrid)
            for (int i = lane; i < 32 * W; i += 32) X[s].ring[(size_t)i * NVO] = qnan;   // every cell starts empty
    }
    __syncwarp();
    uint32_t iters = 0;

    auto flush_half = [&](double *ring, double *partial, long long f0, long long f1) {   // (adaptive_stream's, per stream)
        __syncwarp();
        const int cnt = (int)(f1 - f0);
        const int w0 = (int)(f0 % R), k0 = (int)(f0 % nt);
        double *row = V.wsum + (size_t)lane * A.nslots;
        for (int cc = 0; cc < cnt; cc += 32) {
            const int o = cc + lane;
            const bool mine = o < cnt;
            const int w = (w0 + o) % R, kq = (k0 + o) % nt;
            for (int sl = 0; sl < A.nslots; ++sl) row[sl] = 0.0;
            unsigned nv = 0;
            constexpr int RB = 8;
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += RB) {
                double xa[RB][NVO];
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const double *cell = ring + ((size_t)w * 32 + (r0 + q)) * NVO;
                    if (mine) {
                        if constexpr (NVO == 4) {
                            const double2 a = __ldcg(reinterpret_cast<const double2 *>(cell)), b = __ldcg(reinterpret_cast<const double2 *>(cell + 2));
                            xa[q][0] = a.x; xa[q][1] = a.y; xa[q][2] = b.x; xa[q][3] = b.y;
                        } else {
                            const double2 a = __ldcg(reinterpret_cast<const double2 *>(cell));
                            xa[q][0] = a.x; xa[q][1] = a.y;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < NVO; ++i) xa[q][i] = qnan;
                    }
                }
#pragma unroll
                for (int q = 0; q < RB; ++q)
                    if (mine) ring[((size_t)w * 32 + (r0 + q)) * NVO] = qnan;   // the cells are free again
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const bool valid = (xa[q][0] == xa[q][0]);
                    nv += valid ? 1u : 0u;
                    jit_reduce<NVO, true>(V, &A.smp, A.u64, xa[q][0], xa[q][1], NVO == 4 ? xa[q][NVO - 2] : 0.0, NVO == 4 ? xa[q][NVO - 1] : 0.0, valid, kq, 0, valid);
                }
            }
            if (mine) {
                for (int sl = 0; sl < A.nslots; ++sl) atomicAdd(&partial[(size_t)kq * A.nslots + sl], row[sl]);
                if (nv) atomicAdd(&nvalid_g[kq], (unsigned long long)nv);
            }
        }
        __syncwarp();
    };
    auto emit_state = [&](const double (&uo)[NV], bool valid, double *cell) {
        if (!valid) return;
        double xo[NVO];
        from_chart<SYS>(uo, xo);
        if constexpr (NVO == 4) {
            *reinterpret_cast<double2 *>(cell) = make_double2(xo[0], xo[1]);
            *reinterpret_cast<double2 *>(cell + 2) = make_double2(xo[2], xo[3]);
        } else
            *reinterpret_cast<double2 *>(cell) = make_double2(xo[0], xo[1]);
    };

#pragma unroll 1
    while (true) {
        // ================= service part (cold), stream by stream =================
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            Stream &x = X[s];
            if (x.done) continue;   // (warp-uniform)
            if (!x.have && x.j < A.N) {   // draw the next trajectory of the stream
                x.idx = (uint64_t)x.j;
                x.start_k = 0;
                if (A.redo) {
                    const FailRec r = A.redo[x.j];
                    x.idx = r.idx;
                    x.start_k = r.k;
                }
                x.acc_steps += T[s].nacc;
                x.rej_steps += T[s].nrej;
                T[s].status = ST_OK;
                {
                    double xi[NVO];
                    draw_initial<NVO>(A.smp, x.idx, A.attempt, xi);
                    to_chart<SYS>(xi, T[s].u);
                }
                T[s].begin(A.S);
                x.have = true;
                x.k = 0;
                x.tstop = A.S.ts[0];
                x.budget = (int)(A.S.maxiters > 0x7fffffffll ? 0x7fffffffll : A.S.maxiters);
            }
            if (x.have && T[s].status != ST_OK) {   // failures end a trajectory: report, skip the rest of its outputs
                bool queued = false;
                if (T[s].status != ST_MAXITERS) {
                    atomicAdd(&A.u64[CNT_FAILED], 1ull);
                    if (A.fail_out && A.smp.kind != DTWA_SAMPLER_HOST_ARRAY && A.smp.kind != DTWA_SAMPLER_DEVICE_ARRAY) {
                        const unsigned pos = atomicAdd(A.fail_count, 1u);
                        if (pos < A.fail_cap) {
                            FailRec r;
                            r.idx = x.idx;
                            r.k = x.k;
                            r.pad = 0;
                            A.fail_out[pos] = r;
                            queued = true;
                        }
                    }
                }
                if (!queued) atomicAdd(&A.u64[CNT_UNFINISHED], 1ull);
                x.g += (long long)(nt - x.k);
                x.wslot = (int)(((long long)x.wslot + (nt - x.k)) % R);
                x.k = nt;
            }
            if (x.have && x.k >= nt) {
                x.have = false;
                x.j += stride;
            }
            {   // flush what every lane of the stream has passed
                const bool retired = !x.have && x.j >= A.N;
                const long long ahead = retired ? (long long)0x7fffffff : x.g - x.flushed;
                const int amin = __reduce_min_sync(0xffffffffu, (int)(ahead > 0x7fffffff ? 0x7fffffff : ahead));
                if (amin == 0x7fffffff) {   // the stream has ended: flush its tail
                    const long long gend = __reduce_max_sync(0xffffffffu, (int)(x.g - x.flushed)) + x.flushed;
                    while (x.flushed < gend) {
                        const long long f1 = (x.flushed + H < gend) ? x.flushed + H : gend;
                        flush_half(x.ring, x.partial, x.flushed, f1);
                        x.flushed = f1;
                    }
                    x.done = true;
                } else if (amin >= H) {
                    flush_half(x.ring, x.partial, x.flushed, x.flushed + H);
                    x.flushed += H;
                }
            }
        }
        if (X[0].done && X[1].done) break;
        if (__any_sync(0xffffffffu, (!X[0].done && !X[0].have && X[0].j < A.N) || (!X[1].done && !X[1].have && X[1].j < A.N))) continue;

        // ================= hot loop: one attempted step of each stream per lane and iteration =================
        int ahead0 = X[0].have ? (int)(X[0].g - X[0].flushed) : 0, ahead1 = X[1].have ? (int)(X[1].g - X[1].flushed) : 0;
        const bool live0 = !X[0].done, live1 = !X[1].done;
#pragma unroll 1
        while (true) {
            // leave when a lane needs service or a half of a ring can be flushed
            const bool service = (X[0].have && (T[0].status != ST_OK || X[0].k >= nt)) || (X[1].have && (T[1].status != ST_OK || X[1].k >= nt));
            const bool behind0 = X[0].have && ahead0 < H, behind1 = X[1].have && ahead1 < H;
            const unsigned vote = __reduce_or_sync(0xffffffffu, (service ? 1u : 0u) | (behind0 ? 2u : 0u) | (behind1 ? 4u : 0u));
            if ((vote & 1u) || (live0 && !(vote & 2u)) || (live1 && !(vote & 4u))) break;
            ++iters;
            const bool step0 = X[0].have && ahead0 < R, step1 = X[1].have && ahead1 < R;
            const bool want0 = step0 && T[0].ad.t < X[0].tstop, want1 = step1 && T[1].ad.t < X[1].tstop;
            if (want0 || want1) {
                bool act[2] = {want0 && X[0].budget > 0, want1 && X[1].budget > 0};
                bool forced[2], last[2], clipped[2];
                double h[2], habs_eff[2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if ((s ? want1 : want0) && !act[s]) T[s].status = ST_MAXITERS;
                    X[s].budget -= act[s] ? 1 : 0;
                    forced[s] = T[s].ad.habs < A.S.dtmin;
                    habs_eff[s] = forced[s] ? A.S.dtmin : T[s].ad.habs;
                    last[s] = (T[s].ad.t + habs_eff[s] >= X[s].tstop);
                    clipped[s] = last[s] && ((X[s].tstop - T[s].ad.t) < habs_eff[s]);
                    h[s] = last[s] ? X[s].tstop - T[s].ad.t : habs_eff[s];
                    if (!act[s]) h[s] = 0.0;
                }
                double yn[2][NV], kn[2][NV], err[2];
                float le2[2];
                Adaptive<SYS, ALG>::trial2(A.S.par, T[0].ad, T[1].ad, T[0].u, T[1].u, h[0], h[1], A.S.tol, yn[0], yn[1], kn[0], kn[1], err[0], err[1], le2[0],
                                           le2[1]);
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    // the decisions of Adaptive::attempt, as selects
                    const bool bad = state_bad<SYS>(yn[s]) || !(err[s] == err[s]) || isinf(err[s]);
#ifdef DTWA_LE_FROM_FACTORS
                    const float le = le2[s];
#else
                    const float le = lg2_approx(fminf(fmaxf((float)err[s], 1e-30f), 1e30f));
#endif
                    const float e_rej = ex2_approx(-c.b1 * le);
                    const float e_acc = ex2_approx(-fmaf(c.b1, le, c.nb2 * T[s].ad.lqold));
                    float fac_acc = fminf(c.fac_max, fmaxf(c.fac_min_acc, c.g * e_acc));
                    if (T[s].ad.rejected) fac_acc = fminf(fac_acc, c.post_reject_cap);
                    const float fac_rej = fmaxf(c.fac_min_rej, c.g * e_rej);
                    const bool acc = act[s] && !bad && (forced[s] || err[s] < c.acc_thr);
                    const bool fail = act[s] && bad && forced[s];
                    double hn = h[s] * (double)(acc ? fac_acc : fac_rej);
                    if (bad) hn = h[s] * c.dom_shrink;
                    if (acc && clipped[s] && hn < habs_eff[s]) hn = habs_eff[s];
                    if (acc && !clipped[s]) T[s].ad.lqold = fmaxf(le, c.lqoldinit);
                    if (act[s] && !fail) T[s].ad.habs = hn;
                    if (act[s] && fail) T[s].ad.habs = habs_eff[s];
                    if (act[s]) T[s].ad.rejected = !acc;
                    if (acc) {
                        T[s].ad.t = last[s] ? X[s].tstop : T[s].ad.t + h[s];
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            T[s].u[i] = yn[s][i];
                            T[s].ad.k0[i] = kn[s][i];
                        }
                    }
                    if (fail) T[s].status = ST_DOMAIN;
                    T[s].nacc += acc ? 1u : 0u;
                    T[s].nrej += (act[s] && !acc && !fail) ? 1u : 0u;
                }
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const bool stepping = s ? step1 : step0;
                const bool reached = stepping && T[s].status == ST_OK && !(T[s].ad.t < X[s].tstop);   // this stream emits output k now
                if (reached) {
                    emit_state(T[s].u, X[s].k >= X[s].start_k, X[s].ring + ((size_t)X[s].wslot * 32 + lane) * NVO);
                    ++X[s].k;
                    if (s) ++ahead1; else ++ahead0;
                    X[s].wslot = (X[s].wslot + 1 == R) ? 0 : X[s].wslot + 1;
                    if (X[s].k < nt) X[s].tstop = A.S.ts[X[s].k];
                }
            }
        }
        if (X[0].have) X[0].g = X[0].flushed + ahead0;
        if (X[1].have) X[1].g = X[1].flushed + ahead1;
    }
    // ---- step counters: one atomic per warp ----
    unsigned long long a = X[0].acc_steps + X[1].acc_steps + T[0].nacc + T[1].nacc, r = X[0].rej_steps + X[1].rej_steps + T[0].nrej + T[1].nrej;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        r += __shfl_xor_sync(0xffffffffu, r, off);
    }
    if (lane == 0) {
        atomicAdd(&A.u64[CNT_ACC], a);
        if (r) atomicAdd(&A.u64[CNT_REJ], r);
        atomicAdd(&A.u64[CNT_LANE], 64ull * iters);
    }
}
#endif

template <int SYS, int ALG, bool JIT>
__device__ __forceinline__ void ensemble_body(const EnsArgs &A)
{
    constexpr int NV = Sys<SYS>::NV, NVO = Sys<SYS>::NVO;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool FIXED = is_fixed<ALG>::value;
    constexpr bool PERLANE = !FIXED;  // adaptive: one staging row per thread, no warp-level reduction
    const int nwarps = blockDim.x >> 5;
    const int rows = PERLANE ? (int)blockDim.x : nwarps;   // rows of the shared-memory staging (adaptive kernels: one row of sums per LANE, used by the ring flush)
#ifdef DTWA_PAIR_UNIT
    const int vcta = blockIdx.x;
#else
    const int vcta = PERLANE ? (int)(blockIdx.x * nwarps + (threadIdx.x >> 5)) : (int)blockIdx.x;   // adaptive kernels: the warp's stream number
#endif
    const int lane = threadIdx.x & 31;
    const int W = A.window;
    const int Ws = PERLANE ? 1 : W;   // output times staged in SHARED memory (the adaptive kernels' ring lives in global memory)

    // ---- carve shared memory and stage the reducer program ----
    VmShared V;
    {
        unsigned char *p = smem_raw;
        double *consts = reinterpret_cast<double *>(p);
        p += sizeof(double) * A.ncst;
        RangeD *ranges = reinterpret_cast<RangeD *>(p);
        p += sizeof(RangeD) * A.nrange;
        HistDesc *hists = reinterpret_cast<HistDesc *>(p);
        p += sizeof(HistDesc) * A.nhist;
        V.wsum = reinterpret_cast<double *>(p);
        p += sizeof(double) * (size_t)rows * Ws * (A.nslots > 0 ? A.nslots : 1);
        V.stage = reinterpret_cast<unsigned *>(p);
        p += sizeof(unsigned) * (size_t)Ws * A.stage_bins;
        V.nvalid = reinterpret_cast<unsigned *>(p);
        p += sizeof(unsigned) * Ws;
        V.stage2d = reinterpret_cast<unsigned *>(p);
        p += sizeof(unsigned) * A.stage2d_bins;
        VmIns *prog = reinterpret_cast<VmIns *>(p);
        for (int i = threadIdx.x; i < A.ncst; i += blockDim.x) consts[i] = A.consts[i];
        for (int i = threadIdx.x; i < A.nrange; i += blockDim.x) ranges[i] = A.ranges[i];
        for (int i = threadIdx.x; i < A.nhist; i += blockDim.x) hists[i] = A.hists[i];
        for (int i = threadIdx.x; i < A.nins; i += blockDim.x) prog[i] = A.prog[i];
        for (int i = threadIdx.x; i < Ws * A.stage_bins; i += blockDim.x) V.stage[i] = 0u;
        for (int i = threadIdx.x; i < Ws; i += blockDim.x) V.nvalid[i] = 0u;
        for (int i = threadIdx.x; i < A.stage2d_bins; i += blockDim.x) V.stage2d[i] = 0u;
#ifndef DTWA_PAIR_UNIT   // (the two-stream kernels initialise the rings of their streams themselves)
        if constexpr (PERLANE) {   // the warp's ring of raw states in global memory: every cell starts empty (NaN marker)
            if (vcta < A.vgrid) {
                double *ring = A.ring + (size_t)vcta * 32 * W * Sys<SYS>::NVO;
                for (int i = lane; i < 32 * W; i += 32) ring[(size_t)i * Sys<SYS>::NVO] = __longlong_as_double(0x7ff8000000000000ll);
            }
        }
#endif
        V.consts = consts;
        V.ranges = ranges;
        V.hists = hists;
        V.prog = prog;
        V.nslots = A.nslots;
        V.stage_bins = A.stage_bins;
        V.nwarps = nwarps;
        V.window = Ws;
    }
    __syncthreads();

#if defined(DTWA_DIAG_SMID)
    if (threadIdx.x == 0) {   // diagnostic: which SM hosts this CTA, and when
        unsigned smid;
        unsigned long long t0;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        printf("CTA %d sm %u t %llu\n", (int)blockIdx.x, smid, t0);
    }
#endif
#if defined(DTWA_DIAG_WARPID)
    if (lane == 0) {   // diagnostic: which SM and which warp slot (slot % 4 = scheduler) hosts this warp
        unsigned smid, wid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        printf("WARP %d sm %u slot %u\n", vcta, smid, wid);
    }
#endif
    const int nt = A.S.nt;
    double *my_partial = A.partial + (size_t)vcta * nt * A.nslots;
    unsigned long long *nvalid_g = A.u64 + CNT_N;

    // small 2-D histograms (all times in one matrix) live in a per-CTA shared-memory tile: one shared-memory atomic per sample,
    // one global RED per non-empty cell when the tile is flushed (at the end of the CTA's life, and often enough on the way
    // that no 32-bit cell can overflow)
    auto flush_tiles = [&]() {
        __syncthreads();
        for (int hI = 0; hI < A.nhist; ++hI) {
            const HistDesc h = V.hists[hI];
            if (!h.two_d || h.staged_off < 0) continue;
            const int cells = h.n0 * h.n1;
            for (int i = threadIdx.x; i < cells; i += blockDim.x) {
                const unsigned c = V.stage2d[h.staged_off + i];
                if (c) {
                    atomicAdd(&A.u64[h.base + i], (unsigned long long)c);
                    V.stage2d[h.staged_off + i] = 0u;
                }
            }
        }
        __syncthreads();
    };
    const long long tile_flush_every = A.stage2d_bins > 0 ? max(1ll, (1ll << 31) / ((long long)blockDim.x * max(1, A.S.nt))) : 0;
    long long batches_done = 0;

    if constexpr (!FIXED) {
#ifdef DTWA_PAIR_UNIT
        adaptive_stream_pair<SYS, ALG>(A, V, nvalid_g);
#else
        if (vcta >= A.vgrid) return;   // (a CTA's last warps when the number of streams is not a multiple of DTWA_ADAPT_WARPS; no CTA barrier follows)
        adaptive_stream<SYS, ALG, JIT>(A, V, my_partial, nvalid_g, A.ring + (size_t)vcta * 32 * W * Sys<SYS>::NVO, vcta);
#endif
    } else {
    // ---- fixed step: persistent loop over batches of blockDim.x trajectories ----
    for (long long batch = blockIdx.x; batch * (long long)blockDim.x < A.N; batch += gridDim.x) {
        if (A.stage2d_bins > 0 && ++batches_done % tile_flush_every == 0) flush_tiles();
        const long long j = batch * (long long)blockDim.x + threadIdx.x;
        const bool active = j < A.N;
        Traj<SYS, ALG> T;
        uint64_t idx = (uint64_t)j;
        int start_k = 0;
        T.status = active ? ST_OK : ST_INACTIVE;
#pragma unroll
        for (int i = 0; i < NV; ++i) T.u[i] = 0.0;
        if (active) {
            if (A.redo) {
                const FailRec r = A.redo[j];
                idx = r.idx;
                start_k = r.k;
            }
            double x[NVO];
            draw_initial<NVO>(A.smp, idx, A.attempt, x);
            to_chart<SYS>(x, T.u);
        }
        T.begin(A.S);
        bool failed_reported = false;
        uint32_t warp_iters = 0;  // max over lanes of step iterations, summed over intervals

        auto report_failure = [&](int k) {
            failed_reported = true;
            bool queued = false;
            if (T.status != ST_MAXITERS) {
                atomicAdd(&A.u64[CNT_FAILED], 1ull);
                const bool can_redo = A.fail_out && A.smp.kind != DTWA_SAMPLER_HOST_ARRAY && A.smp.kind != DTWA_SAMPLER_DEVICE_ARRAY;
                if (can_redo) {
                    const unsigned pos = atomicAdd(A.fail_count, 1u);
                    if (pos < A.fail_cap) {
                        FailRec r;
                        r.idx = idx;
                        r.k = k;
                        r.pad = 0;
                        A.fail_out[pos] = r;
                        queued = true;
                    }
                }
            }
            if (!queued) atomicAdd(&A.u64[CNT_UNFINISHED], 1ull);
        };
        auto reduce_at = [&](bool valid, int k, int w) {   // observables + reducers into the window staging
            double xo[NVO];
            from_chart<SYS>(T.u, xo);   // the observables' coordinates (the identity except in the Bloch chart)
            if constexpr (JIT)
                jit_reduce<NVO, PERLANE>(V, &A.smp, A.u64, xo[0], xo[1], NVO == 4 ? xo[NVO - 2] : 0.0, NVO == 4 ? xo[NVO - 1] : 0.0,
                                         valid, k, w);
            else
                vm_run<NVO, PERLANE>(V, &A.smp, A.u64, xo[0], xo[1], NVO == 4 ? xo[NVO - 2] : 0.0, NVO == 4 ? xo[NVO - 1] : 0.0,
                                     valid, k, w, nullptr);
        };

        if constexpr (FIXED) {
        for (int k0 = 0; k0 < nt; k0 += W) {
            const int nw = min(W, nt - k0);
            // all lanes take the same steps: advance together, reduce with warp shuffles
            for (int w = 0; w < nw; ++w) {
                const int k = k0 + w;
                warp_iters += T.advance(A.S, k);
                // one rarely taken branch per output time: the chart test runs unconditionally (cheap), everything else
                // only for lanes that are not (or no longer) healthy
                if (state_bad<SYS>(T.u) || T.status != ST_OK) {
                    if (T.status == ST_OK) T.status = ST_DOMAIN;
                    if (T.status != ST_INACTIVE && !failed_reported) report_failure(k);
                }
                const bool valid = (T.status == ST_OK) && (k >= start_k);
                reduce_at(valid, k, w);
                // fixed step: V.nvalid counts the threads that do NOT contribute (failed, inactive tail,
                // before start_k), so the common all-valid case costs one vote and no shared-memory atomic
                const unsigned inv = __ballot_sync(0xffffffffu, !valid);
                if (inv && lane == 0) atomicAdd(&V.nvalid[w], (unsigned)__popc(inv));
            }
            // ---- once per window: CTA-wide flush in a fixed order (deterministic sums) ----
            __syncthreads();
            for (int i = threadIdx.x; i < nw * A.nslots; i += blockDim.x) {
                const int w = i / A.nslots, sl = i - w * A.nslots;
                double s = 0.0;
                for (int rw = 0; rw < rows; ++rw) s = __dadd_rn(s, V.wsum[((size_t)rw * W + w) * A.nslots + sl]);
                // only this CTA ever adds to its partial row: RED in program order => reproducible
                atomicAdd(&my_partial[(size_t)(k0 + w) * A.nslots + sl], s);
            }
            for (int w = threadIdx.x; w < nw; w += blockDim.x) {
                const unsigned c = blockDim.x - V.nvalid[w];
                if (c) atomicAdd(&nvalid_g[k0 + w], (unsigned long long)c);
                V.nvalid[w] = 0u;
            }
            for (int hI = 0; hI < A.nhist; ++hI) {
                const HistDesc h = V.hists[hI];
                if (h.staged_off < 0 || h.two_d) continue;   // (2-D tiles are flushed by flush_tiles)
                for (int i = threadIdx.x; i < nw * h.n0; i += blockDim.x) {
                    const int w = i / h.n0, b = i - w * h.n0;
                    unsigned *cell = V.stage + (size_t)w * A.stage_bins + h.staged_off + b;
                    const unsigned c = *cell;
                    if (c) {
                        atomicAdd(&A.u64[h.base + (long long)(k0 + w) * h.tstride + b], (unsigned long long)c);
                        *cell = 0u;
                    }
                }
            }
            __syncthreads();
        }
        }
        // ---- step counters: one atomic per warp per batch ----
        const unsigned long long a = __reduce_add_sync(0xffffffffu, T.nacc);
        const unsigned long long r = __reduce_add_sync(0xffffffffu, T.nrej);
        if (lane == 0) {
            atomicAdd(&A.u64[CNT_ACC], a);
            if (r) atomicAdd(&A.u64[CNT_REJ], r);
            atomicAdd(&A.u64[CNT_LANE], 32ull * warp_iters);
        }
    }
    if (A.stage2d_bins > 0) flush_tiles();
    }  // fixed step
}

template <int SYS, int ALG>
__global__ void __launch_bounds__(ens_launch<ALG>::max_threads, ens_launch<ALG>::min_blocks) ensemble_kernel(const EnsArgs A)
{
    ensemble_body<SYS, ALG, false>(A);
}

#ifndef DTWA_JIT_UNIT  // everything below is only needed in the ahead-of-time build
// partial[grid][n] -> out[n], CTA order fixed => run-to-run reproducible sums
__global__ void finalize_sums_kernel(const double *__restrict__ partial, int grid, long long n, double *__restrict__ out)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int c = 0; c < grid; ++c) s = __dadd_rn(s, partial[(size_t)c * n + i]);
    out[i] = s;
}

// ------------------------------------------------------------------------------------------------
// K3: dense trajectories
template <int SYS, int ALG>
__global__ void __launch_bounds__(128) integrate_kernel(const SolverArgs S, const double *__restrict__ u0, long long n,
                                                        double *__restrict__ out, int32_t *__restrict__ status,
                                                        long long *__restrict__ nsteps)
{
    constexpr int NV = Sys<SYS>::NV, NVO = Sys<SYS>::NVO;
    const long long j0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = j0 < n;
    const long long j = live ? j0 : n - 1;  // tail lanes shadow the last trajectory: warp-wide ops need every lane
    Traj<SYS, ALG> T;
    T.status = ST_OK;
    {
        double x[NVO];
#pragma unroll
        for (int i = 0; i < NVO; ++i) x[i] = u0[j * NVO + i];
        to_chart<SYS>(x, T.u);
    }
    T.begin(S);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    auto emit = [&](int k) {
        if (T.status == ST_OK && state_bad<SYS>(T.u)) T.status = ST_DOMAIN;
        if (live) {
            double xo[NVO];
            from_chart<SYS>(T.u, xo);
            double *o = out + ((size_t)j * S.nt + k) * NVO;
#pragma unroll
            for (int i = 0; i < NVO; ++i) o[i] = (T.status == ST_OK) ? xo[i] : qnan;
        }
    };
    if constexpr (is_fixed<ALG>::value) {
        for (int k = 0; k < S.nt; ++k) {
            T.advance(S, k);
            emit(k);
        }
    } else {
        // flat loop: one attempted step per lane and iteration; a lane that has reached its output time stores the state and
        // moves on to the next one without waiting for the others (nested "for each output time: while (t < t_k) step"
        // loops make every lane wait for the slowest one at every output time)
        int k = 0;
        if (!S.interp) {
#pragma unroll 1
        while (__any_sync(0xffffffffu, k < S.nt)) {
            if (k < S.nt) {
                const double tstop = S.ts[k];
                if (T.status == ST_OK && T.ad.t < tstop) {
                    if ((long long)(T.nacc + T.nrej) >= S.maxiters) {
                        T.status = ST_MAXITERS;
                    } else {
                        bool accepted;
                        if (T.ad.attempt(S.par, T.u, tstop, S.tol, S.dtmin, accepted, S.ctrl))
                            T.status = ST_DOMAIN;
                        else if (accepted)
                            ++T.nacc;
                        else
                            ++T.nrej;
                    }
                }
                if (T.status != ST_OK || !(T.ad.t < tstop)) {
                    emit(k);
                    ++k;
                }
            }
        }
        } else {
        // DTWA_OUTPUT_HERMITE: same step sequence as the ensemble kernel's (clip at the maxout-th output time ahead, see
        // adaptive_stream: maxout = min(32, half of a ring of min(256, nt) outputs)), outputs from the Hermite interpolant
        const int wfull = S.nt < 256 ? S.nt : 256;
        const int maxout = min(INTERP_MAX_OUT, wfull >= 2 ? wfull / 2 : 1);
#pragma unroll 1
        while (__any_sync(0xffffffffu, k < S.nt)) {
            if (k < S.nt) {
                bool accepted = false;
                double yprev[NV], fprev[NV];
                if (T.status == ST_OK && T.ad.t < S.ts[k]) {
                    if ((long long)(T.nacc + T.nrej) >= S.maxiters) {
                        T.status = ST_MAXITERS;
                    } else {
                        const double tlim = S.ts[min(k + maxout - 1, S.nt - 1)];
                        if (T.ad.template attempt<true>(S.par, T.u, tlim, S.tol, S.dtmin, accepted, S.ctrl, yprev, fprev))
                            T.status = ST_DOMAIN;
                        else if (accepted)
                            ++T.nacc;
                        else
                            ++T.nrej;
                    }
                }
                while (k < S.nt && (T.status != ST_OK || !(T.ad.t < S.ts[k]))) {
                    const double tk = S.ts[k];
                    if (T.status == ST_OK && accepted && tk < T.ad.t) {
                        double uo[NV], keep[NV];
                        const double theta = (tk - (T.ad.t - T.ad.hlast)) / T.ad.hlast;
                        hermite_interpolant<NV>(yprev, fprev, T.u, T.ad.k0, T.ad.hlast, theta, uo);
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            keep[i] = T.u[i];
                            T.u[i] = uo[i];
                        }
                        emit(k);   // (reports the interpolated state; a bad one ends the trajectory here)
#pragma unroll
                        for (int i = 0; i < NV; ++i) T.u[i] = keep[i];
                    } else
                        emit(k);
                    ++k;
                }
            }
        }
        }
    }
    if (live) {
        status[j] = T.status;
        if (nsteps) nsteps[j] = (long long)T.nacc + (long long)T.nrej;
    }
}

// ------------------------------------------------------------------------------------------------
// K5: trajectories with their fundamental matrix (src/ClassicalSystems.jl:110-128).  SYS is one of the
// variational systems; GROUP adjacent threads share one initial condition (dtwa_device.cuh).
//   out_u [n][nt][NVB], out_fm[n][nt][NVB*NVB] column-major (Julia's layout): thread c writes x_c and column c.
template <int SYS, int ALG>
__global__ void __launch_bounds__(128) integrate_fm_kernel(const SolverArgs S, const double *__restrict__ u0,
                                                           const double *__restrict__ fm0, long long n,
                                                           double *__restrict__ out_u, double *__restrict__ out_fm,
                                                           int32_t *__restrict__ status, long long *__restrict__ nsteps)
{
    constexpr int NVB = Sys<SYS>::NVB, G = Sys<SYS>::GROUP;
    const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c = (int)(gid % G);
    const long long j0 = gid / G;
    const bool live = j0 < n;
    const long long j = live ? j0 : n - 1;  // tail groups shadow the last initial condition
    Traj<SYS, ALG> T;
    T.status = ST_OK;
    T.u[0] = u0[j * NVB + c];
#pragma unroll
    for (int i = 0; i < NVB; ++i) T.u[1 + i] = fm0 ? fm0[(j * NVB + c) * NVB + i] : (i == c ? 1.0 : 0.0);
    T.begin(S);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    auto emit = [&](int k) {
        if (T.status == ST_OK && state_bad<SYS>(T.u)) T.status = ST_DOMAIN;
        if (live) {
            const bool ok = T.status == ST_OK;
            out_u[((size_t)j * S.nt + k) * NVB + c] = ok ? T.u[0] : qnan;
            double *o = out_fm + (((size_t)j * S.nt + k) * NVB + c) * NVB;
#pragma unroll
            for (int i = 0; i < NVB; ++i) o[i] = ok ? T.u[1 + i] : qnan;
        }
    };
    if constexpr (is_fixed<ALG>::value) {
        for (int k = 0; k < S.nt; ++k) {
            T.advance(S, k);
            emit(k);
        }
    } else {   // flat loop, as in integrate_kernel; every decision is uniform inside a group
        int k = 0;
#pragma unroll 1
        while (__any_sync(0xffffffffu, k < S.nt)) {
            if (k < S.nt) {
                const double tstop = S.ts[k];
                if (T.status == ST_OK && T.ad.t < tstop) {
                    if ((long long)(T.nacc + T.nrej) >= S.maxiters) {
                        T.status = ST_MAXITERS;
                    } else {
                        bool accepted;
                        if (T.ad.attempt(S.par, T.u, tstop, S.tol, S.dtmin, accepted, S.ctrl))
                            T.status = ST_DOMAIN;
                        else if (accepted)
                            ++T.nacc;
                        else
                            ++T.nrej;
                    }
                }
                if (T.status != ST_OK || !(T.ad.t < tstop)) {
                    emit(k);
                    ++k;
                }
            }
        }
    }
    if (live && c == 0) {
        status[j] = T.status;
        if (nsteps) nsteps[j] = (long long)T.nacc + (long long)T.nrej;
    }
}

// K6: Lyapunov spectra, one group per initial condition -- src/ClassicalSystems.jl:183-224, the
// algorithm of Fig. 3.7 of Parker & Chua as the reference codes it: integrate (x, Phi = U) over an
// interval t, modified Gram-Schmidt of the columns against the NEW u_j (j < i), sums_i += log|v_i|,
// lambda_i = sums_i / (k t), and -- inside the loop over i -- stop as soon as
// |lambda_old - lambda| < lreltol |lambda| + ltol (so the returned vector may mix entries of
// iteration k and k-1, exactly like the reference).  Each interval is a fresh `integrate` call
// (new starting step).  status: 0 ok, DTWA_TRAJ_* if the integration failed, DTWA_LYAP_MAXITERS.
struct LyapArgs {
    double t_interval, ltol, lreltol;
    int32_t maxiters, pad;
};
// Orbits need very different numbers of rounds (4 .. maxiters), so the groups of a persistent grid draw their
// next initial condition from a global counter (`next`, zeroed by the host) instead of owning a fixed one.
#ifndef DTWA_LYAP_MIN_BLOCKS
#define DTWA_LYAP_MIN_BLOCKS 2   // flat loop: 255 registers without spills at 8 warps/SM beat 168 registers with 350 B of spills at 12 (382 vs 402 ms, 256^2 map)
#endif
template <int SYS, int ALG>
__global__ void __launch_bounds__(128, DTWA_LYAP_MIN_BLOCKS) lyapunov_kernel(const SolverArgs S, const LyapArgs L, const double *__restrict__ u0,
                                                       long long n, double *__restrict__ lambda_out,
                                                       double *__restrict__ x_out, int32_t *__restrict__ iters_out,
                                                       int32_t *__restrict__ status, long long *__restrict__ nsteps,
                                                       unsigned long long *__restrict__ next)
{
    constexpr int NVB = Sys<SYS>::NVB, G = Sys<SYS>::GROUP;
    const int c = (int)(threadIdx.x % G);
    const unsigned base = (threadIdx.x & 31u) & ~(unsigned)(G - 1);
    const unsigned mask = ((1u << G) - 1u) << base;
    // Gram-Schmidt of the columns, lambda_i and the convergence test of round k (src/ClassicalSystems.jl:199-218), group-wide
    auto renormalise = [&](double (&w)[NVB], double &sums, double (&lam)[NVB], const double (&lam_old)[NVB], int k) -> bool {
        bool done = false;
        const double kt = __dmul_rn((double)k, L.t_interval);
#pragma unroll
        for (int i = 0; i < NVB; ++i) {
            if (done) break;
#pragma unroll
            for (int jj = 0; jj < i; ++jj) {
                double uj[NVB];
#pragma unroll
                for (int r = 0; r < NVB; ++r) uj[r] = __shfl_sync(mask, w[r], (int)base + jj);
                if (c == i) {
                    double d = 0.0;
#pragma unroll
                    for (int r = 0; r < NVB; ++r) d = __dadd_rn(d, __dmul_rn(w[r], uj[r]));
#pragma unroll
                    for (int r = 0; r < NVB; ++r) w[r] = __dsub_rn(w[r], __dmul_rn(d, uj[r]));
                }
            }
            double li = 0.0;
            if (c == i) {
                double n2 = 0.0;
#pragma unroll
                for (int r = 0; r < NVB; ++r) n2 = __dadd_rn(n2, __dmul_rn(w[r], w[r]));
                const double nrm = __dsqrt_rn(n2);
#pragma unroll
                for (int r = 0; r < NVB; ++r) w[r] = __ddiv_rn(w[r], nrm);
                sums = __dadd_rn(sums, log(nrm));
                li = __ddiv_rn(sums, kt);
            }
            lam[i] = __shfl_sync(mask, li, (int)base + i);
            double e2 = 0.0, l2 = 0.0;
#pragma unroll
            for (int r = 0; r < NVB; ++r) {
                const double dl = __dsub_rn(lam_old[r], lam[r]);
                e2 = __dadd_rn(e2, __dmul_rn(dl, dl));
                l2 = __dadd_rn(l2, __dmul_rn(lam[r], lam[r]));
            }
            if (__dsqrt_rn(e2) < __dadd_rn(__dmul_rn(L.lreltol, __dsqrt_rn(l2)), L.ltol)) done = true;
        }
        return done;
    };
    if constexpr (!is_fixed<ALG>::value) {
        // Adaptive pairs: FLAT loop.  Every iteration each group that is inside an interval attempts ONE step; starting a
        // round (initial step size), ending it (renormalisation, convergence test) and drawing the next orbit are short
        // group-local episodes.  With the loops nested (orbits > rounds > steps) the eight groups of a warp, which need
        // different numbers of steps per interval and of rounds per orbit, executed the step body one after the other:
        // 17.5 active lanes per executed instruction (ncu) instead of ~31.
        bool have = false, exhausted = false, fresh = false;
        long long j = 0, steps = 0;
        double xc = 0.0, w[NVB], lam[NVB], lam_old[NVB], sums = 0.0;
        int k = 0;
        Traj<SYS, ALG> T;
        T.status = ST_OK;
        T.nacc = T.nrej = 0;
#pragma unroll
        for (int i = 0; i < NVB; ++i) w[i] = lam[i] = lam_old[i] = 0.0;
#pragma unroll
        for (int i = 0; i < Sys<SYS>::NV; ++i) T.u[i] = 0.0;
        const double tstop = S.ts[0];
        auto finish = [&](int st, int it) {
            if (x_out) x_out[j * NVB + c] = xc;
            if (c == 0) {
#pragma unroll
                for (int i = 0; i < NVB; ++i) lambda_out[j * NVB + i] = lam[i];
                iters_out[j] = it;
                status[j] = st;
                if (nsteps) nsteps[j] = steps;
            }
            have = false;
        };
#pragma unroll 1
        while (true) {
            if (!have && !exhausted) {   // draw the next orbit (group-uniform)
                unsigned long long jn = 0;
                if (c == 0) jn = atomicAdd(next, 1ull);
                jn = __shfl_sync(mask, jn, (int)base);
                if (jn >= (unsigned long long)n)
                    exhausted = true;
                else {
                    j = (long long)jn;
                    xc = u0[j * NVB + c];
#pragma unroll
                    for (int i = 0; i < NVB; ++i) {
                        w[i] = (i == c) ? 1.0 : 0.0;
                        lam[i] = 0.0;
                    }
                    sums = 0.0;
                    steps = 0;
                    k = 0;
                    have = true;
                    fresh = true;
                }
            }
            if (!__any_sync(0xffffffffu, have)) break;
            if (have && fresh) {   // a round is a fresh `integrate` call: new starting step
                ++k;
#pragma unroll
                for (int i = 0; i < NVB; ++i) lam_old[i] = lam[i];
                T.status = ST_OK;
                T.u[0] = xc;
#pragma unroll
                for (int i = 0; i < NVB; ++i) T.u[1 + i] = w[i];
                T.begin(S);
                fresh = false;
            }
            if (have && T.status == ST_OK && T.ad.t < tstop) {   // one attempted step
                if ((long long)(T.nacc + T.nrej) >= S.maxiters) {
                    T.status = ST_MAXITERS;
                } else {
                    bool accepted;
                    if (T.ad.attempt(S.par, T.u, tstop, S.tol, S.dtmin, accepted, S.ctrl))
                        T.status = ST_DOMAIN;
                    else if (accepted)
                        ++T.nacc;
                    else
                        ++T.nrej;
                }
            }
            if (have && (T.status != ST_OK || !(T.ad.t < tstop))) {   // the interval has ended
                if (T.status == ST_OK && state_bad<SYS>(T.u)) T.status = ST_DOMAIN;
                steps += (long long)T.nacc + (long long)T.nrej;
                if (T.status != ST_OK) {
                    finish(T.status, k);
                } else {
                    xc = T.u[0];
#pragma unroll
                    for (int i = 0; i < NVB; ++i) w[i] = T.u[1 + i];
                    const bool done = renormalise(w, sums, lam, lam_old, k);
                    if (done)
                        finish(ST_OK, k);
                    else if (k >= L.maxiters)
                        finish(DTWA_LYAP_MAXITERS, k);
                    else
                        fresh = true;
                }
            }
        }
    } else {
#pragma unroll 1
    for (;;) {
    unsigned long long jn = 0;
    if (c == 0) jn = atomicAdd(next, 1ull);
    jn = __shfl_sync(mask, jn, (int)base);
    if (jn >= (unsigned long long)n) break;
    const long long j = (long long)jn;
    const bool live = true;
    double xc = u0[j * NVB + c], w[NVB], lam[NVB], lam_old[NVB];   // this thread's component of x, its column of Phi
    double sums = 0.0;
#pragma unroll
    for (int i = 0; i < NVB; ++i) {
        w[i] = (i == c) ? 1.0 : 0.0;
        lam[i] = 0.0;
    }
    int st = ST_OK, it = 0;
    long long steps = 0;
    bool done = false;
#pragma unroll 1
    for (int k = 1; k <= L.maxiters && !done && st == ST_OK; ++k) {
        it = k;
#pragma unroll
        for (int i = 0; i < NVB; ++i) lam_old[i] = lam[i];
        Traj<SYS, ALG> T;
        T.status = ST_OK;
        T.u[0] = xc;
#pragma unroll
        for (int i = 0; i < NVB; ++i) T.u[1 + i] = w[i];
        T.begin(S);
        T.advance(S, 0);
        if (T.status == ST_OK && state_bad<SYS>(T.u)) T.status = ST_DOMAIN;
        steps += (long long)T.nacc + (long long)T.nrej;
        if (T.status != ST_OK) {
            st = T.status;
            break;
        }
        xc = T.u[0];
#pragma unroll
        for (int i = 0; i < NVB; ++i) w[i] = T.u[1 + i];
        done = renormalise(w, sums, lam, lam_old, k);
    }
    if (st == ST_OK && !done) st = DTWA_LYAP_MAXITERS;
    if (live && x_out) x_out[j * NVB + c] = xc;
    if (live && c == 0) {
#pragma unroll
        for (int i = 0; i < NVB; ++i) lambda_out[j * NVB + i] = lam[i];
        iters_out[j] = it;
        status[j] = st;
        if (nsteps) nsteps[j] = steps;
    }
    }  // work loop
    }  // fixed step
}

// K7: surfaces of section -- SURVEY 8f rank 3 (docs/src/ClassicalDickeExamples.md:82-109: integrate with a
// DiffEq ContinuousCallback on p = 0, save_positions=(false,false)).  One thread per initial condition
// integrates to t_end and records every crossing of g(u) = coef . u - offset = 0 (direction: 0 both,
// +1 g increasing, -1 decreasing).  Where the reference searches a root of its interpolant, the crossing is
// located to integrator accuracy: starting from the chord estimate, Newton on the crossing time, each
// iteration re-taking the integrator's own step from the start of the crossing step (local error of an
// accepted step); what is left of |g| (rounding level) is removed by one Henon step (classical RK4 in the event variable, dtwa_device.cuh), so
// recorded states satisfy g = 0 to rounding.
// Divergence: a crossing costs ~5 extra steps and happens in ~3 % of the steps of a lane.  A lane therefore
// only NOTES a crossing (step start, times, g values) and keeps stepping; the warp refines all noted
// crossings together when some lane meets its next one (or at the end), so the refinement runs with most
// lanes active instead of one.
//   ev_t[n][cap], ev_u[n][cap][NV], ev_dir[n][cap] (+1/-1), ev_count[n] (keeps counting beyond cap)
struct EventArgs {
    double coef[4];
    double offset, t_end;
    int32_t direction, cap;
};
#ifndef DTWA_EVENTS_MIN_BLOCKS
#define DTWA_EVENTS_MIN_BLOCKS 12   // 168 registers (a few spilled doubles), 12 warps per SM: 10 % faster than the 212 registers ptxas takes unasked
#endif
constexpr int EVENTS_BLOCK = 32;   // one warp per CTA: lanes of a warp share the refinement service, warps share nothing
template <int SYS, int ALG>
__global__ void __launch_bounds__(EVENTS_BLOCK, DTWA_EVENTS_MIN_BLOCKS) events_kernel(const SolverArgs S, const EventArgs E, const double *__restrict__ u0,
                                                     long long n, double *__restrict__ ev_t, double *__restrict__ ev_u,
                                                     int32_t *__restrict__ ev_dir, int32_t *__restrict__ ev_count,
                                                     double *__restrict__ u_end, int32_t *__restrict__ status,
                                                     long long *__restrict__ nsteps)
{
    constexpr int NV = Sys<SYS>::NV;
    constexpr int HSYS = (SYS == DTWA_SYS_DICKE) ? DTWA_SYS_DICKE_HENON : DTWA_SYS_LMG_HENON;
    constexpr bool FIXED = is_fixed<ALG>::value;
    const long long j0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = j0 < n;
    const long long j = live ? j0 : n - 1;
    HenonPar hp;
    static_cast<SysPar &>(hp) = S.par;
#pragma unroll
    for (int i = 0; i < 4; ++i) hp.coef[i] = E.coef[i];
    auto gfun = [&](const double (&u)[NV]) {
        double g = -E.offset;
#pragma unroll
        for (int i = 0; i < NV; ++i) g = fma(E.coef[i], u[i], g);
        return g;
    };
    auto own_step = [&](double (&z)[NV], double th) {   // the integrator's own one-step map
        if constexpr (ALG == DTWA_ALG_RK4) {
            StepTab r;
            const double wf = (SYS == DTWA_SYS_DICKE) ? S.par.b : 1.0;
            r.h = th; r.hh = th / 2.0; r.h3 = th / 3.0; r.h6 = th / 6.0;
            r.hw = r.h * wf; r.hhw = r.hh * wf; r.h3w = r.h3 * wf; r.h6w = r.h6 * wf;
            r.n = 1; r.pad = 0;
            rk4_step<SYS>(S.par, z, r);
        } else {
            rk8_step<SYS>(S.par, z, th);
        }
    };
    // The Henon step has length |g| <= 1e-13 (what Newton left), so a classical RK4 step in the event variable is exact
    // to rounding whatever the integrator; a 12-stage step here would only be one more large body in the instruction cache.
    auto henon = [&](double (&y)[NV + 1], double hg) {
        StepTab r;
        r.h = hg; r.hh = hg / 2.0; r.h3 = hg / 3.0; r.h6 = hg / 6.0;
        r.hw = r.hhw = r.h3w = r.h6w = 0.0;
        r.n = 1; r.pad = 0;
        rk4_step<HSYS>(hp, y, r);
    };
    Traj<SYS, ALG> T;
    T.status = ST_OK;
#pragma unroll
    for (int i = 0; i < NV; ++i) T.u[i] = u0[j * NV + i];
    T.begin(S);
    int count = 0;
    double t = 0.0, g0 = (T.status == ST_OK) ? gfun(T.u) : 0.0;
    // the noted crossing of this lane
    bool pending = false;
    // (step-start state, its time; the bracket [plo, phi] of step fractions -- in time units -- that holds the
    //  crossing, and g at its ends: the whole step for an ordinary crossing, one side of the extremum for a tangency)
    double ps[NV], pts0 = 0.0, plo = 0.0, phi = 0.0, pg0 = 0.0, pg1 = 0.0;
    // The refinement service.  Every integrator body is inlined exactly ONCE in this kernel (the adaptive attempt, the
    // own-step map, the Henon step): a lane's refinement jobs -- its noted crossing, the probe of a tangency at the
    // interpolant's extremum, the one or two crossings of a tangency -- all go through one loop as a small per-lane
    // state machine.  (Five inlined copies of the refinement, as in the first version, made 255 registers with spills
    // and an instruction-cache-bound kernel.)
    enum { SV_IDLE = 0, SV_NEWTON = 1, SV_PROBE = 2, SV_FIN = 3 };
    int mode = SV_IDLE, nit = 0;
    double lo = 0.0, hi = 0.0, glo = 0.0, th = 0.0, g2 = 0.0, gx = 0.0, z[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) z[i] = ps[i] = 0.0;
    auto start_newton = [&]() {   // chord estimate inside the bracket of the noted crossing
        lo = plo;
        hi = phi;
        glo = pg0;
        th = lo + (hi - lo) * (pg0 / (pg0 - pg1));
        if (pg1 == 0.0) th = hi;
        nit = 0;
        mode = SV_NEWTON;
    };
    const StepTab r = FIXED ? S.tab[0] : StepTab();
    int sidx = 0;
    bool done = !(T.status == ST_OK && E.t_end > 0.0);
#pragma unroll 1
    while (true) {
        const bool all_done = !__any_sync(0xffffffffu, !done);
        if (all_done && !__any_sync(0xffffffffu, pending)) break;
        bool cross = false, accepted = false, tang = false;
        double cs[NV], cts0 = 0.0, chs = 0.0, cg1 = 0.0, gd0 = 0.0, thx = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) cs[i] = T.u[i];
        if (!done) {
            if constexpr (!FIXED) {   // dg/dt at the step start, from the FSAL derivative
#pragma unroll
                for (int i = 0; i < NV; ++i) gd0 = fma(E.coef[i], T.ad.k0[i], gd0);
            }
            double t1 = t;
            if constexpr (FIXED) {
                if constexpr (ALG == DTWA_ALG_RK4)
                    rk4_step<SYS>(S.par, T.u, r);
                else
                    rk8_step<SYS>(S.par, T.u, r.h);
                ++T.nacc;
                ++sidx;
                if (state_bad<SYS>(T.u)) {
                    T.status = ST_DOMAIN;
                    done = true;
                } else {
                    accepted = true;
                    t1 = (sidx == r.n) ? E.t_end : r.h * (double)sidx;
                    if (sidx == r.n) done = true;
                }
            } else {
                if ((long long)(T.nacc + T.nrej) >= S.maxiters) {
                    T.status = ST_MAXITERS;
                    done = true;
                } else if (T.ad.attempt(S.par, T.u, E.t_end, S.tol, S.dtmin, accepted, S.ctrl)) {
                    T.status = ST_DOMAIN;
                    done = true;
                    accepted = false;
                } else if (accepted) {
                    ++T.nacc;
                    t1 = T.ad.t;
                    if (!(T.ad.t < E.t_end)) done = true;
                } else
                    ++T.nrej;
            }
            if (accepted) {
                cg1 = gfun(T.u);
                const bool up = g0 < 0.0 && cg1 >= 0.0, down = g0 > 0.0 && cg1 <= 0.0;
                cross = (up && E.direction >= 0) || (down && E.direction <= 0);
                cts0 = t;
                chs = t1 - t;
                t = t1;
                if constexpr (!FIXED) {
                    // Tangency: no sign change over the step, but g turns round inside it (dg/dt changes sign) and the
                    // cubic Hermite interpolant of g dips through zero at its extremum -> two crossings in this step.
                    // (DiffEq samples `interp_points` inside every step for the same purpose.)
                    double gd1 = 0.0;
#pragma unroll
                    for (int i = 0; i < NV; ++i) gd1 = fma(E.coef[i], T.ad.k0[i], gd1);
                    if (!up && !down && g0 * cg1 > 0.0 && gd0 * gd1 < 0.0) {
                        const double d0 = chs * gd0, d1 = chs * gd1, dl = cg1 - g0;
                        const double c2 = 3.0 * dl - 2.0 * d0 - d1, c3 = d0 + d1 - 2.0 * dl;   // g = g0 + d0 x + c2 x^2 + c3 x^3
                        // g'(x) = d0 + 2 c2 x + 3 c3 x^2 = 0: the root in (0, 1)
                        const double qa = 3.0 * c3, qb = 2.0 * c2, disc = qb * qb - 4.0 * qa * d0;
                        double x = -1.0;
                        if (disc >= 0.0) {
                            const double sq = sqrt(disc), q = -0.5 * (qb + (qb >= 0.0 ? sq : -sq));
                            const double x1 = (qa != 0.0) ? q / qa : -1.0, x2 = (q != 0.0) ? d0 / q : -1.0;
                            x = (x1 > 0.0 && x1 < 1.0) ? x1 : x2;
                        }
                        if (x > 0.0 && x < 1.0) {
                            const double gxi = g0 + x * (d0 + x * (c2 + x * c3));
                            if (gxi * g0 < 0.0) {
                                tang = true;
                                thx = x * chs;
                            }
                        }
                    }
                }
            }
        }
        // Noted crossings are refined together when some lane needs its slot again (it crossed again, or it met a
        // tangency, which it handles in time order: noted crossing first), or when every lane has finished.
        if (all_done || __any_sync(0xffffffffu, (cross && pending) || tang)) {
            bool probe_left = tang, second = false;
            if (pending)
                start_newton();
            else if (probe_left) {
                probe_left = false;
                mode = SV_PROBE;
                th = thx;
            }
#pragma unroll 1
            while (true) {
                const bool stepping = mode == SV_NEWTON || mode == SV_PROBE;
                if (__any_sync(0xffffffffu, stepping)) {
                    if (stepping) {   // the integrator's own step from the start of the crossing step
#pragma unroll
                        for (int i = 0; i < NV; ++i) z[i] = (mode == SV_PROBE) ? cs[i] : ps[i];
                        own_step(z, th);
                        g2 = gfun(z);
                        if (mode == SV_PROBE) {   // the integrator's own value at the interpolant's extremum
                            mode = SV_IDLE;
                            if (g2 * g0 < 0.0) {
                                const bool first_up = g0 < 0.0;
                                const bool want1 = (first_up && E.direction >= 0) || (!first_up && E.direction <= 0);
                                const bool want2 = (!first_up && E.direction >= 0) || (first_up && E.direction <= 0);
                                gx = g2;
#pragma unroll
                                for (int i = 0; i < NV; ++i) ps[i] = cs[i];
                                pts0 = cts0;
                                if (want1) {
                                    plo = 0.0; phi = thx; pg0 = g0; pg1 = gx;
                                    second = want2;
                                    start_newton();
                                } else if (want2) {
                                    plo = thx; phi = chs; pg0 = gx; pg1 = cg1;
                                    start_newton();
                                }
                            }
                        } else if (nit == 7 || fabs(g2) <= 1e-13) {
                            mode = SV_FIN;
                        } else {
                            if ((g2 < 0.0) == (glo < 0.0)) {              // keep the bracket
                                lo = th;
                                glo = g2;
                            } else
                                hi = th;
                            double f[NV], gd = 0.0;                       // Newton on the crossing time, bisection when it leaves the bracket
                            Sys<SYS>::rhs(S.par, z, f);
#pragma unroll
                            for (int i = 0; i < NV; ++i) gd = fma(E.coef[i], f[i], gd);
                            double tn = th - g2 / gd;
                            if (!(tn > lo) || !(tn < hi)) tn = 0.5 * (lo + hi);
                            th = tn;
                            ++nit;
                        }
                    }
                    continue;
                }
                // no lane is stepping: the located crossings take their Henon step together
                if (!__any_sync(0xffffffffu, mode == SV_FIN)) break;
                if (mode == SV_FIN) {
                    double y[NV + 1];
#pragma unroll
                    for (int i = 0; i < NV; ++i) y[i] = z[i];
                    y[NV] = pts0 + th;
                    if (g2 != 0.0) {
                        henon(y, -g2);                            // |g2| is at rounding level unless Newton left the step
                        bool fin = true;
#pragma unroll
                        for (int i = 0; i <= NV; ++i) fin = fin && (y[i] == y[i]) && !isinf(y[i]);
                        if (!fin) {
#pragma unroll
                            for (int i = 0; i < NV; ++i) y[i] = z[i];
                            y[NV] = pts0 + th;
                        }
                    }
                    if (count < E.cap && live) {
                        const size_t e = (size_t)j * E.cap + count;
                        ev_t[e] = y[NV];
#pragma unroll
                        for (int i = 0; i < NV; ++i) ev_u[e * NV + i] = y[i];
                        ev_dir[e] = (pg0 < 0.0) ? 1 : -1;
                    }
                    ++count;
                    pending = false;
                    mode = SV_IDLE;
                    if (second) {            // the way back through the section of a tangency
                        second = false;
                        plo = thx; phi = chs; pg0 = gx; pg1 = cg1;
                        start_newton();
                    } else if (probe_left) {
                        probe_left = false;
                        mode = SV_PROBE;
                        th = thx;
                    }
                }
            }
        }
        if (cross) {
            pending = true;
#pragma unroll
            for (int i = 0; i < NV; ++i) ps[i] = cs[i];
            pts0 = cts0;
            plo = 0.0;
            phi = chs;
            pg0 = g0;
            pg1 = cg1;
        }
        if (accepted) g0 = cg1;
    }
    if (live) {
        ev_count[j] = count;
        status[j] = T.status;
#pragma unroll
        for (int i = 0; i < NV; ++i) u_end[j * NV + i] = T.u[i];
        if (nsteps) nsteps[j] = (long long)T.nacc + (long long)T.nrej;
    }
}

// Compaction of the padded event arrays: one warp per initial condition copies its min(count, cap) events to
// [off[i], off[i+1]) (contiguous, coalesced).
__global__ void __launch_bounds__(256) pack_events_kernel(const double *__restrict__ ev_t, const double *__restrict__ ev_u,
                                                          const int32_t *__restrict__ ev_dir, const int64_t *__restrict__ off,
                                                          long long n, long long cap, int nv, double *__restrict__ out_t,
                                                          double *__restrict__ out_u, int32_t *__restrict__ out_dir)
{
    const long long i = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const long long o = off[i], m = off[i + 1] - o;
    for (long long k = lane; k < m; k += 32) {
        out_t[o + k] = ev_t[i * cap + k];
        out_dir[o + k] = ev_dir[i * cap + k];
    }
    for (long long k = lane; k < m * nv; k += 32) out_u[o * nv + k] = ev_u[i * cap * nv + k];
}

// K8: energy-shell quadrature -- SURVEY 8f rank 4, src/EnergyShellProjections.jl:49-96 (the double integral of
// f(x) delta(h_cl(x) - eps) over q and p by Chebyshev-Gauss quadrature in p, Eq. (8) of Pilatowsky-Cameo et al.,
// Nat. Commun. 2021).  One warp per (Q, P) cell; the lanes stride over the 2 n (or n) nodes, evaluate the
// nexpr classical expressions with the observable VM at [Q, q_+-(Q,P,p,eps), P, p] and the warp reduces.
//   out[ncell][nexpr] (NaN where the reference returns `nonvalue`), valid[ncell], nodes[ncell]
struct ShellArgs {
    double w0, w, g, eps, p_res;
    int32_t onlyqroot;   // 0: both roots of q, +1 / -1: only q_+ / q_-
    int32_t nexpr;
    const VmIns *prog;   // nexpr programs `<expr> STORE END`, back to back
    const int32_t *prog_off;
    const double *consts;
};
constexpr int SHELL_MAX_EXPR = 8;
__global__ void __launch_bounds__(128) shell_quadrature_kernel(const ShellArgs A, const double *__restrict__ QP, long long ncell,
                                                               double *__restrict__ out, int32_t *__restrict__ valid,
                                                               int32_t *__restrict__ nodes)
{
    const long long cell = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cell >= ncell) return;
    const double Q = QP[2 * cell], P = QP[2 * cell + 1];
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    const double R = Q * Q + P * P, r = 1.0 - R / 4.0;
    // src/EnergyShellProjections.jl:57-72
    const double Aq = 4.0 * A.g * A.g * (Q * Q) * r - A.w * A.w0 * (P * P + Q * Q - 2.0) + 2.0 * A.w * A.eps;
    const double pplus = sqrt(Aq) / A.w - 1e-10;
    const bool ok = !(R > 4.0) && !(Aq < 0.0) && !(pplus < 0.0);
    if (!ok) {
        if (lane == 0) {
            for (int e = 0; e < A.nexpr; ++e) out[cell * A.nexpr + e] = qnan;
            valid[cell] = 0;
            nodes[cell] = 0;
        }
        return;
    }
    const int n = (int)(2.0 * floor(pplus / A.p_res) + 1.0);
    const int nroots = A.onlyqroot == 0 ? 2 : 1;
    const double wgt = DTWA_PI / ((double)n * A.w);
    const double sr = sqrt(r);
    double acc[SHELL_MAX_EXPR];
#pragma unroll
    for (int e = 0; e < SHELL_MAX_EXPR; ++e) acc[e] = 0.0;
    VmShared V = {};
    V.consts = A.consts;
    V.window = 1;
    for (int idx = lane; idx < n * nroots; idx += 32) {
        const int root = idx / n, k = idx - root * n + 1;
        const double sgn = (A.onlyqroot != 0) ? (double)A.onlyqroot : (root == 0 ? 1.0 : -1.0);
        const double p = pplus * cos(DTWA_PI * (double)(2 * k - 1) / (double)(2 * n));
        // q_of_eps, src/ClassicalDicke.jl:224-228,255-273
        const double disc = 4.0 * A.g * A.g * (Q * Q) * r - (p * p) * (A.w * A.w) - A.w * A.w0 * (P * P + Q * Q - 2.0) + 2.0 * A.w * A.eps;
        const double q = (-2.0 * A.g * Q * sr + sgn * sqrt(disc)) / A.w;   // disc < 0 -> NaN (the reference raises)
#pragma unroll 1
        for (int e = 0; e < A.nexpr; ++e) {
            V.prog = A.prog + A.prog_off[e];
            double v = 0.0;
            vm_run<4, false>(V, nullptr, nullptr, Q, q, P, p, true, 0, 0, &v);
            acc[e] += wgt * v;
        }
    }
#pragma unroll
    for (int e = 0; e < SHELL_MAX_EXPR; ++e) {
        if (e < A.nexpr) {
            double sum = acc[e];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, off));
            if (lane == 0) out[cell * A.nexpr + e] = sum;
        }
    }
    if (lane == 0) {
        valid[cell] = 1;
        nodes[cell] = n * nroots;
    }
}

// ------------------------------------------------------------------------------------------------
// hooks
template <int NV>
__global__ void sample_kernel(const SamplerPar sp, long long n, uint32_t attempt, double *__restrict__ u)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    double x[NV];
    draw_initial<NV>(sp, (uint64_t)j, attempt, x);
#pragma unroll
    for (int i = 0; i < NV; ++i) u[j * NV + i] = x[i];
}

template <int NV>
__global__ void pdf_kernel(const SamplerPar sp, const double *__restrict__ u, long long n, double *__restrict__ w)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    double x[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) x[i] = u[j * NV + i];
    w[j] = pdf_eval<NV>(sp, x);
}

template <int SYS>
__global__ void hamiltonian_kernel(const double p0, const double p1, const double p2, const double *__restrict__ u,
                                   long long n, double *__restrict__ h)
{
    constexpr int NV = Sys<SYS>::NV;
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double par[3] = {p0, p1, p2};
    double x[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) x[i] = u[j * NV + i];
    h[j] = Sys<SYS>::hamiltonian(par, x);
}

// expression evaluation through the same VM as the ensemble kernel (program: <expr> OP_STORE OP_END)
template <int NV>
__global__ void eval_kernel(const VmIns *__restrict__ prog, const double *__restrict__ consts, const double *__restrict__ u,
                            long long n, double *__restrict__ v)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    VmShared V = {};
    V.consts = consts;
    V.prog = prog;
    V.window = 1;
    double x[4] = {0, 0, 0, 0};
    if (j < n) {
#pragma unroll
        for (int i = 0; i < NV; ++i) x[i] = u[j * NV + i];
    }
    double outv = 0.0;
    vm_run<NV, false>(V, nullptr, nullptr, x[0], x[1], x[2], x[3], j < n, 0, 0, &outv);
    if (j < n) v[j] = outv;
}

__global__ void bin_kernel(const double *__restrict__ v, long long n, const RangeD r, long long *__restrict__ idx)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    idx[j] = scale_to_index(v[j], r);
}

// K0: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// K0c: CHAINS independent DFMA chains per thread, one warp per CTA, any number of CTAs per SM -- how much of the FP64 pipe
// a given (instruction-level parallelism x warps per scheduler) can keep busy.  The adaptive kernels run 3 warps per
// scheduler with 4 independent chains in their stage sums; this kernel reproduces exactly that shape without anything else.
template <int CHAINS, int NREG>
__global__ void __launch_bounds__(32) fp64_chains_kernel(double *out, const double *in, int iters, double a, double b)
{
    // NREG distinct vector-register operands per DFMA: 1 = x*a + b (a, b warp-uniform: K0's form), 2 = x*a + y (the form of a
    // stage sum: coefficient uniform, stage value and accumulator in registers), 3 = x*z + y
    double x[CHAINS], y[CHAINS], z[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x[c] = threadIdx.x * 1e-3 + c;
        y[c] = in[(threadIdx.x + c) & 63] * 1e-7;
        z[c] = 1.0 - in[(threadIdx.x + 2 * c + 1) & 63] * 1e-6;
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 64 / CHAINS; ++r) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if constexpr (NREG == 1)
                    x[c] = fma(x[c], a, b);
                else if constexpr (NREG == 2)
                    x[c] = fma(x[c], a, y[c]);
                else
                    x[c] = fma(x[c], z[c], y[c]);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// K0d: 8 DFMA chains per thread (K0's form) with ALU other instructions woven in per group of 8 DFMAs -- does an
// instruction of another pipe issue "for free" in the cycle the FP64 pipe cannot accept a second DFMA, or does it take an
// issue slot away from the DFMAs?  KIND 0: integer adds (ALU pipe), 1: FP32 FMAs (FMA pipe), 2: MUFU (XU pipe), 3: IMAD (integer
// multiply-add: FMA pipe -- the pipe ptxas also sends half of its register moves to, as IMAD.MOV), 4: LOP3 (ALU pipe).
template <int ALU, int KIND>
__global__ void __launch_bounds__(32) fp64_mix_kernel(double *out, int iters, double a, double b)
{
    double x[8];
    unsigned u[4];
    float f[4];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x * 1e-3 + c;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        u[c] = threadIdx.x + c;
        f[c] = 1.0f + threadIdx.x * 1e-3f + c;
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                x[c] = fma(x[c], a, b);
                // ALU other instructions per 8 DFMAs, spread evenly
#pragma unroll
                for (int m = 0; m < ((c + 1) * ALU) / 8 - (c * ALU) / 8; ++m) {
                    const int k = (c + m) & 3;
                    if constexpr (KIND == 0)
                        asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(i));
                    else if constexpr (KIND == 1)
                        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[k]) : "f"(1.0000001f));
                    else if constexpr (KIND == 2)
                        asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[k]));
                    else if constexpr (KIND == 3)   // IMAD: integer multiply-add, issued to the FMA pipe like FFMA
                        asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(u[k]) : "r"(i | 3));
                    else                            // LOP3 (ALU pipe)
                        asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[k]) : "r"(i));
                }
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c];
    s += (double)(u[0] ^ u[1] ^ u[2] ^ u[3]) + (double)(f[0] + f[1] + f[2] + f[3]);
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// K0b: the same 8 chains with three distinct REGISTER operands per DFMA (what an ODE right-hand side
// looks like; K0's multiplier/addend are warp-uniform).  Reported next to K0 as the practical ceiling.
__global__ void __launch_bounds__(256) fp64_peak_regs_kernel(double *out, const double *in, int iters)
{
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = in[(threadIdx.x + i) & 63] + threadIdx.x * 1e-3;
        y[i] = in[(threadIdx.x + 8 + i) & 63] * 0.999;
        z[i] = in[(threadIdx.x + 16 + i) & 63] * 1e-7;
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], y[i], z[(i + r) & 7]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#endif  // DTWA_JIT_UNIT

}  // namespace dtwa
