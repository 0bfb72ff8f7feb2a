// dtwa_device.cuh -- device-side building blocks of libdicketwa (sm_100a, FP64, one trajectory per thread).
//
// Everything a trajectory needs lives in registers: state, RK stages, step controller, Philox
// counters.  Floating-point sequences are written with explicit fma()/__dmul_rn()/__dadd_rn() so
// that every kernel instantiating a stepper executes the same instruction sequence and therefore
// produces bit-identical trajectories (dtwa_integrate vs dtwa_ensemble).
//
// Reference behaviour restated (paths relative to the reference checkout):
//   equations of motion   src/ClassicalDicke.jl:11-17, src/ClassicalLMG.jl:8-11
//   domain                src/ClassicalDicke.jl:18-20, src/ClassicalLMG.jl:12-14
//   Hamiltonians          src/ClassicalDicke.jl:80-88, src/ClassicalLMG.jl:63-70
//   samplers / pdfs       src/TWA.jl:689-764, src/PhaseSpaces.jl:16,34,49,63
//   scale_to_index        src/TWA.jl:53-64
#pragma once
// (under NVRTC -- the run-time specialisation of the reducers, dtwa_jit.hpp -- these three names are
//  served from in-memory headers: fixed-width typedefs, an empty cuda_runtime.h, the embedded C header)
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/dicketwa.h"

namespace dtwa {

// ------------------------------------------------------------------------------------------------
// parameters (sign of `integate_backwards` folded in: every term of both vector fields is linear in
// exactly one parameter, so sigma*f(u; p) == f(u; sigma*p))
struct SysPar {
    double a, b, c, d;  // Dicke: w0, w, g, -      LMG: Omega, xi, xi/2, 2*xi+Omega
};

// Two independent values side by side: the arithmetic type of Adaptive::trial2, which integrates two trajectories of one lane
// with every operation issued for the first and then for the second -- the source order that makes ptxas interleave the two
// dependency chains (given the stage code of one and then of the other it leaves them one after the other).
struct D2 {
    double a, b;
};
template <class T>
__device__ __forceinline__ T bc(double c);   // a constant as T
template <>
__device__ __forceinline__ double bc<double>(double c) { return c; }
template <>
__device__ __forceinline__ D2 bc<D2>(double c) { return D2{c, c}; }
__device__ __forceinline__ double dm(double x, double y) { return __dmul_rn(x, y); }
__device__ __forceinline__ D2 dm(D2 x, D2 y) { return D2{__dmul_rn(x.a, y.a), __dmul_rn(x.b, y.b)}; }
__device__ __forceinline__ D2 dm(double x, D2 y) { return D2{__dmul_rn(x, y.a), __dmul_rn(x, y.b)}; }
__device__ __forceinline__ D2 dm(D2 x, double y) { return D2{__dmul_rn(x.a, y), __dmul_rn(x.b, y)}; }
__device__ __forceinline__ double ng(double x) { return -x; }
__device__ __forceinline__ D2 ng(D2 x) { return D2{-x.a, -x.b}; }
__host__ __device__ __forceinline__ double fma(double x, double y, double z) { return ::fma(x, y, z); }   // (the overloads below hide ::fma)
__device__ __forceinline__ D2 fma(D2 x, D2 y, D2 z) { return D2{fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)}; }
__device__ __forceinline__ D2 fma(double x, D2 y, D2 z) { return D2{fma(x, y.a, z.a), fma(x, y.b, z.b)}; }
__device__ __forceinline__ D2 fma(D2 x, D2 y, double z) { return D2{fma(x.a, y.a, z), fma(x.b, y.b, z)}; }
__device__ __forceinline__ D2 fma(double x, D2 y, double z) { return D2{fma(x, y.a, z), fma(x, y.b, z)}; }

// 1/sqrt(x) and sqrt(x) for x > 0 together: MUFU.RSQ64H seed (rel. err ~2^-22) + one third-order
// polish, six FP64 instructions.  No special cases: x <= 0 or non-finite yields NaN/Inf, which the
// domain checks catch.  No instruction reads three distinct registers (a DFMA with three distinct
// register operands needs three register-file cycles against two issue cycles, see
// fp64_peak_regs_kernel) and sqrt(x) does not wait for 1/sqrt(x).
__device__ __forceinline__ void rsqrt_sqrt_pair(double x, double &is, double &s)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double t = __dmul_rn(x, y0);           // ~ sqrt(x)
    const double e = fma(-t, y0, 1.0);           // 1 - x*y0^2
    const double p = fma(0.375, e, 0.5);
    const double ep = __dmul_rn(e, p);           // e/2 + 3/8 e^2
    is = fma(y0, ep, y0);
    s = fma(t, ep, t);
}

__device__ __forceinline__ void rsqrt_sqrt_pair(D2 x, D2 &is, D2 &s)   // (the same operations, a then b)
{
    double ya, yb;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ya) : "d"(x.a));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yb) : "d"(x.b));
    const D2 y0{ya, yb};
    const D2 t = dm(x, y0);
    const D2 e = fma(ng(t), y0, 1.0);
    const D2 p = fma(0.375, e, 0.5);
    const D2 ep = dm(e, p);
    is = fma(y0, ep, y0);
    s = fma(t, ep, t);
}

// 1/x for the error norms of the step controllers: MUFU.RCP64H seed + one Newton step, relative error
// ~1e-13 (the accept/reject threshold err < 1 does not need more), 2 FP64 instructions instead of ~9.
__device__ __forceinline__ double rcp_fast(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    return fma(y0, fma(-x, y0, 1.0), y0);
}
// sqrt(x) for x >= 0 through the same polished rsqrt (0 stays 0)
__device__ __forceinline__ double sqrt_fast(double x)
{
    double is, s;
    rsqrt_sqrt_pair(x, is, s);
    return x > 0.0 ? s : 0.0;
}

template <int SYS>
struct Sys;

template <>
struct Sys<DTWA_SYS_DICKE> {
    static constexpr int NV = 4;
    static constexpr int NVO = 4;                         // phase-space coordinates the samplers / observables see
    static constexpr int QI = 0, PI = 2;
    static constexpr int NVB = 4, GROUP = 1, NORM_N = 4;  // see the variational systems below
    // 18 FP64-pipe instructions + 1 MUFU.  FOLDQ: du[1] is returned as p (the factor w lives in the
    // caller's step constants), 17 instructions.  DTWA_W_IS_ONE / DTWA_W_IS_MINUS_ONE (defined by the
    // run-time specialisation when |w| == 1; x*1.0 == x exactly, so results do not change): 16.  DTWA_G_IS_ONE /
    // DTWA_G_IS_MINUS_ONE (|g| == 1, the same argument for g*q): 15.
    template <bool FOLDQ = false, class T = double>
    __device__ __forceinline__ static void rhs(const SysPar &p, const T (&u)[4], T (&du)[4])
    {
        const T Q = u[0], q = u[1], P = u[2], pp = u[3];
        T r2 = fma(ng(P), P, 4.0);
        r2 = fma(ng(Q), Q, r2);            // 4 - P^2 - Q^2
        T is, s;
        rsqrt_sqrt_pair(r2, is, s);        // 1/s and s
#if defined(DTWA_G_IS_ONE)
        const T gq = q;
#elif defined(DTWA_G_IS_MINUS_ONE)
        const T gq = ng(q);
#else
        const T gq = dm(p.c, q);
#endif
        const T gqQ = dm(gq, Q);
        const T e = fma(gqQ, is, -p.a);           // g q Q / s - w0
        du[0] = dm(ng(P), e);                     // P w0 - g P q Q / s
        if constexpr (FOLDQ)
            du[1] = pp;
        else {
#if defined(DTWA_W_IS_ONE)
            du[1] = pp;
#elif defined(DTWA_W_IS_MINUS_ONE)
            du[1] = ng(pp);
#else
            du[1] = dm(pp, p.b);                  // p w
#endif
        }
        const T gqs = dm(gq, s);
        du[2] = fma(Q, e, ng(gqs));               // g q Q^2/s - g q s - Q w0
        const T Qs = dm(Q, s);
#if defined(DTWA_W_IS_ONE)
        const T qw = q;
#elif defined(DTWA_W_IS_MINUS_ONE)
        const T qw = ng(q);
#else
        const T qw = dm(q, p.b);
#endif
        du[3] = fma(-p.c, Qs, ng(qw));            // -g Q s - q w
    }
    __device__ __forceinline__ static double hamiltonian(const double *par, const double *u)
    {
        // reference operation order, no contraction
        const double w0 = par[0], w = par[1], g = par[2];
        const double Q = u[0], q = u[1], P = u[2], p = u[3];
        const double QQ = __dmul_rn(Q, Q), PP = __dmul_rn(P, P);
        const double R = __dadd_rn(QQ, PP);
        const double r = __dsub_rn(1.0, __ddiv_rn(R, 4.0));
        double h = __dmul_rn(__ddiv_rn(w0, 2.0), R);
        h = __dadd_rn(h, __dmul_rn(__ddiv_rn(w, 2.0), __dadd_rn(__dmul_rn(q, q), __dmul_rn(p, p))));
        h = __dadd_rn(h, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, g), q), Q), __dsqrt_rn(r)));
        return __dsub_rn(h, w0);
    }
};

template <>
struct Sys<DTWA_SYS_LMG> {
    static constexpr int NV = 2;
    static constexpr int NVO = 2;
    static constexpr int QI = 0, PI = 1;
    static constexpr int NVB = 2, GROUP = 1, NORM_N = 2;
    // 7 FP64-pipe instructions
    template <bool FOLDQ = false, class T = double>
    __device__ __forceinline__ static void rhs(const SysPar &p, const T (&u)[2], T (&du)[2])
    {
        const T Q = u[0], P = u[1];
        const T Q2 = dm(Q, Q);
        const T P2 = dm(P, P);
        du[0] = dm(P, fma(-p.c, Q2, p.a));               // P (Om - xi Q^2/2)
        const T in = fma(p.c, P2, -p.d);                 // xi P^2/2 - (2 xi + Om)
        du[1] = dm(Q, fma(p.b, Q2, in));                 // Q(...) + xi Q^3
    }
    __device__ __forceinline__ static double hamiltonian(const double *par, const double *u)
    {
        const double Om = par[0], xi = par[1];
        const double Q = u[0], P = u[1];
        const double QQ = __dmul_rn(Q, Q), PP = __dmul_rn(P, P);
        double a = __ddiv_rn(__dmul_rn(Om, __dadd_rn(QQ, PP)), 2.0);
        double b = __dsub_rn(QQ, __ddiv_rn(__dmul_rn(PP, QQ), 4.0));
        b = __dsub_rn(b, __ddiv_rn(__dmul_rn(QQ, QQ), 4.0));   // Q^4 as (Q*Q)*(Q*Q)
        return __dsub_rn(__dadd_rn(a, __dmul_rn(xi, b)), Om);
    }
};

// ------------------------------------------------------------------------------------------------
// Opt-in chart for the Dicke model (dtwa_solver.chart = DTWA_CHART_BLOCH): the spin as a Cartesian unit vector
//   jx = Q s/2,  jy = -P s/2,  jz = (Q^2 + P^2)/2 - 1,   s = sqrt(4 - Q^2 - P^2)        (src/TWA.jl:816-826, src/PhaseSpaces.jl:49,63)
// in which Hamilton's equations of src/ClassicalDicke.jl:11-17 become bilinear (chain rule, h = w0 jz + w (q^2+p^2)/2 + 2 g q jx):
//   djx = -w0 jy,   djy = w0 jx - 2 g q jz,   djz = 2 g q jy,   dq = w p,   dp = -w q - 2 g jx.
// No square root, no division, no chart singularity at Q^2 + P^2 -> 4 (where the (Q,P) field grows like 1/s and an adaptive
// integrator legitimately needs 10^4-10^5 steps of ~dtmin): 8 FP64 instructions per right-hand side instead of 19 + MUFU.
// The SAME differential equation in other coordinates: exact solutions coincide, numerical ones differ by the truncation
// error of the method, so this chart leaves the reference's step sequence -- it is never the default.
// State y = [jx, jy, jz, q, p]; samplers and observables keep seeing (Q, q, P, p) through to_chart / from_chart.
constexpr int DTWA_SYS_DICKE_BLOCH = 6;
template <>
struct Sys<DTWA_SYS_DICKE_BLOCH> {
    static constexpr int NV = 5;
    static constexpr int NVO = 4;
    static constexpr int QI = 0, PI = 1;   // (unused: no chart boundary)
    static constexpr int NVB = 5, GROUP = 1, NORM_N = 5;
    // p.a = w0, p.b = w, p.c = g, p.d = 2 g (sign of integate_backwards folded into all of them)
    template <bool FOLDQ = false, class T = double>
    __device__ __forceinline__ static void rhs(const SysPar &p, const T (&y)[5], T (&dy)[5])
    {
        const T jx = y[0], jy = y[1], jz = y[2], q = y[3], pp = y[4];
        const T gq2 = dm(p.d, q);                          // 2 g q
        dy[0] = dm(-p.a, jy);
        dy[1] = fma(p.a, jx, ng(dm(gq2, jz)));
        dy[2] = dm(gq2, jy);
#if defined(DTWA_W_IS_ONE)
        dy[3] = pp;
        dy[4] = fma(-p.d, jx, ng(q));
#elif defined(DTWA_W_IS_MINUS_ONE)
        dy[3] = ng(pp);
        dy[4] = fma(-p.d, jx, q);
#else
        dy[3] = dm(p.b, pp);
        dy[4] = fma(-p.d, jx, ng(dm(p.b, q)));
#endif
    }
};

// phase-space point (the samplers', observables' and callers' coordinates) <-> integration state
template <int SYS>
__device__ __forceinline__ void to_chart(const double (&x)[Sys<SYS>::NVO], double (&y)[Sys<SYS>::NV])
{
    if constexpr (SYS == DTWA_SYS_DICKE_BLOCH) {
        const double Q = x[0], P = x[2];
        const double R = fma(P, P, __dmul_rn(Q, Q));
        const double hs = __dmul_rn(0.5, sqrt(4.0 - R));   // NaN outside the disk: caught by the finiteness test
        y[0] = __dmul_rn(Q, hs);
        y[1] = -__dmul_rn(P, hs);
        y[2] = fma(0.5, R, -1.0);
        y[3] = x[1];
        y[4] = x[3];
    } else {
#pragma unroll
        for (int i = 0; i < Sys<SYS>::NV; ++i) y[i] = x[i];
    }
}
template <int SYS>
__device__ __forceinline__ void from_chart(const double (&y)[Sys<SYS>::NV], double (&x)[Sys<SYS>::NVO])
{
    if constexpr (SYS == DTWA_SYS_DICKE_BLOCH) {
        // (Q, P) = (jx, -jy) sqrt(2 / (1 - jz)), which gives Q^2 + P^2 = 2 (1 + jz) on the unit sphere
        const double c = 1.0 - y[2];
        double is, sq;
        rsqrt_sqrt_pair(__dmul_rn(0.5, c), is, sq);        // 1 / sqrt((1 - jz)/2)
        x[0] = __dmul_rn(y[0], is);
        x[2] = -__dmul_rn(y[1], is);
        x[1] = y[3];
        x[3] = y[4];
    } else {
#pragma unroll
        for (int i = 0; i < Sys<SYS>::NV; ++i) x[i] = y[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Variational systems (src/ClassicalSystems.jl:110-128, `get_fundamental_matrix=true`):
//   d/dt (x, Phi) = (f(x), J(x) Phi),  J = df/dx  (the reference takes `step(system).jac` from
//   ParameterizedFunctions' symbolic differentiation; here it is written out).
// The columns of Phi are independent given x(t), so a GROUP of NVB adjacent threads integrates one
// initial condition: thread c carries COMPONENT c of x and column c of Phi -- state y = [x_c | Phi_{.,c}],
// 1 + NVB doubles per thread instead of NVB + NVB^2, everything stays in registers.  Every right-hand side
// gathers x from the group (NVB shuffles) and evaluates f and J once per thread; the stage sums -- most of
// the work of the 12-stage pair -- are not replicated.  (The first version carried all of x in every thread:
// 8 instead of 5 components per thread in every stage sum, 234 registers instead of 170.)
// The error norm of the adaptive pairs is a group-wide sum over the NORM_N = NVB + NVB^2 components,
// combined with a butterfly of shuffles inside the group, so all threads of a group take the same steps.
constexpr int DTWA_SYS_DICKE_VAR = 2, DTWA_SYS_LMG_VAR = 3;
template <int G>
__device__ __forceinline__ unsigned group_base()
{
    return (threadIdx.x & 31u) & ~(unsigned)(G - 1);
}
template <int G>
__device__ __forceinline__ unsigned group_mask()
{
    return ((1u << G) - 1u) << group_base<G>();
}

template <>
struct Sys<DTWA_SYS_DICKE_VAR> {
    static constexpr int NV = 5;
    static constexpr int QI = 0, PI = 2;   // group lanes that hold Q and P
    static constexpr int NVB = 4, GROUP = 4, NORM_N = 20;
    __device__ __forceinline__ static void rhs(const SysPar &p, const double (&y)[5], double (&dy)[5])
    {
        const unsigned gb = group_base<4>(), gm = group_mask<4>();
        const int lane_c = (int)(threadIdx.x & 3u);
        const double Q = __shfl_sync(gm, y[0], (int)gb + 0), q = __shfl_sync(gm, y[0], (int)gb + 1);
        const double P = __shfl_sync(gm, y[0], (int)gb + 2), pp = __shfl_sync(gm, y[0], (int)gb + 3);
        const double a = y[1], b = y[2], c = y[3], d = y[4];   // perturbation of (Q, q, P, p)
        double r2 = fma(-P, P, 4.0);
        r2 = fma(-Q, Q, r2);
        double is, s;
        rsqrt_sqrt_pair(r2, is, s);
        const double gq = __dmul_rn(p.c, q);
        const double f = __dmul_rn(__dmul_rn(gq, Q), is);       // g q Q / s
        const double e = f - p.a;                               // f - w0
        const double f0 = __dmul_rn(-P, e);
        const double f1 = __dmul_rn(pp, p.b);
        const double f2 = fma(Q, e, -__dmul_rn(gq, s));
        const double f3 = fma(-p.c, __dmul_rn(Q, s), -__dmul_rn(q, p.b));
        dy[0] = (lane_c & 2) ? ((lane_c & 1) ? f3 : f2) : ((lane_c & 1) ? f1 : f0);   // this thread's component of f(x)
        // df = f_Q a + f_q b + f_P c,  ds = -(Q a + P c)/s
        const double is2 = __dmul_rn(is, is);
        const double fQ = __dmul_rn(__dmul_rn(gq, is), fma(__dmul_rn(Q, Q), is2, 1.0));   // g q (1/s + Q^2/s^3)
        const double fq = __dmul_rn(__dmul_rn(p.c, Q), is);                               // g Q / s
        const double fP = __dmul_rn(__dmul_rn(f, P), is2);                                // g q Q P / s^3
        const double df = fma(fQ, a, fma(fq, b, __dmul_rn(fP, c)));
        const double ds = -__dmul_rn(fma(Q, a, __dmul_rn(P, c)), is);
        dy[1] = fma(-P, df, -__dmul_rn(c, e));                                            // c (w0 - f) - P df
        dy[2] = __dmul_rn(d, p.b);
        dy[3] = fma(a, e, fma(Q, df, -__dmul_rn(p.c, fma(b, s, __dmul_rn(q, ds)))));      // a (f - w0) + Q df - g (b s + q ds)
        dy[4] = fma(-p.c, fma(a, s, __dmul_rn(Q, ds)), -__dmul_rn(b, p.b));               // -g (a s + Q ds) - w b
    }
};

template <>
struct Sys<DTWA_SYS_LMG_VAR> {
    static constexpr int NV = 3;
    static constexpr int QI = 0, PI = 1;   // group lanes that hold Q and P
    static constexpr int NVB = 2, GROUP = 2, NORM_N = 6;
    __device__ __forceinline__ static void rhs(const SysPar &p, const double (&y)[3], double (&dy)[3])
    {
        const unsigned gb = group_base<2>(), gm = group_mask<2>();
        const double Q = __shfl_sync(gm, y[0], (int)gb + 0), P = __shfl_sync(gm, y[0], (int)gb + 1);
        const double a = y[1], c = y[2];
        const double Q2 = __dmul_rn(Q, Q), P2 = __dmul_rn(P, P);
        const double A = fma(-p.c, Q2, p.a);                  // Om - xi Q^2/2   = d(dQ)/dP
        const double in = fma(p.c, P2, -p.d);                 // xi P^2/2 - (2 xi + Om)
        dy[0] = (threadIdx.x & 1u) ? __dmul_rn(Q, fma(p.b, Q2, in)) : __dmul_rn(P, A);   // this thread's component of f(x)
        const double xQP = __dmul_rn(p.b, __dmul_rn(Q, P));   // xi Q P
        const double B = fma(__dmul_rn(3.0, p.b), Q2, in);    // d(dP)/dQ
        dy[1] = fma(A, c, -__dmul_rn(xQP, a));
        dy[2] = fma(B, a, __dmul_rn(xQP, c));
    }
};

// ------------------------------------------------------------------------------------------------
// Henon's trick for surfaces of section (M. Henon, Physica D 5 (1982) 412): with the event function
// g(u) = coef . u - offset as the independent variable, du/dg = f(u) / g', dt/dg = 1 / g',
// g' = coef . f(u), ONE Runge-Kutta step of length -g(u1) from the end of the step that crossed the
// section lands on it (g = 0 to rounding).  State: y = [u (NVB) | t].
struct HenonPar : SysPar {
    double coef[4];
};
template <int BASE>
struct HenonSys {
    static constexpr int NVB = Sys<BASE>::NV, NV = NVB + 1;
    static constexpr int QI = Sys<BASE>::QI, PI = Sys<BASE>::PI, GROUP = 1, NORM_N = NV;
    __device__ __forceinline__ static void rhs(const SysPar &p0, const double (&y)[NV], double (&dy)[NV])
    {
        const HenonPar &p = static_cast<const HenonPar &>(p0);
        double u[NVB], f[NVB];
#pragma unroll
        for (int i = 0; i < NVB; ++i) u[i] = y[i];
        Sys<BASE>::rhs(p, u, f);
        double gd = 0.0;
#pragma unroll
        for (int i = 0; i < NVB; ++i) gd = fma(p.coef[i], f[i], gd);
        const double inv = 1.0 / gd;
#pragma unroll
        for (int i = 0; i < NVB; ++i) dy[i] = f[i] * inv;
        dy[NVB] = inv;
    }
};
constexpr int DTWA_SYS_DICKE_HENON = 4, DTWA_SYS_LMG_HENON = 5;
template <>
struct Sys<DTWA_SYS_DICKE_HENON> : HenonSys<DTWA_SYS_DICKE> {};
template <>
struct Sys<DTWA_SYS_LMG_HENON> : HenonSys<DTWA_SYS_LMG> {};

// does component i of this thread enter the group-wide norms?  (always: no component is replicated in a group)
template <int SYS>
__device__ __forceinline__ bool norm_counted(int)
{
    return true;
}
// sum over the threads of a group; every thread of the group receives the same bits
template <int SYS>
__device__ __forceinline__ double group_sum(double s)
{
    if constexpr (Sys<SYS>::GROUP > 1) {
        constexpr int G = Sys<SYS>::GROUP;
        const unsigned mask = group_mask<G>();
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) s = __dadd_rn(s, __shfl_xor_sync(mask, s, off));
    }
    return s;
}
template <int SYS>
__device__ __forceinline__ double group_meansq(const double (&x)[Sys<SYS>::NV])
{
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < Sys<SYS>::NV; ++i)
        if (norm_counted<SYS>(i)) s = fma(x[i], x[i], s);
    s = group_sum<SYS>(s);
    return s * (1.0 / (double)Sys<SYS>::NORM_N);
}
template <int SYS>
__device__ __forceinline__ double group_rms(const double (&x)[Sys<SYS>::NV])
{
    return sqrt_fast(group_meansq<SYS>(x));
}

template <int SYS>
__device__ __forceinline__ bool state_bad(const double (&u)[Sys<SYS>::NV])
{
    // non-finite anywhere, or outside the chart 4 - P^2 - Q^2 > 0 (un-fused, src/ClassicalDicke.jl:19).
    // A non-finite Q or P makes r2 NaN or -Inf; the other components are tested on their exponent
    // bits (integer pipe -- this runs at every output time, next to a saturated FP64 pipe).
    if constexpr (Sys<SYS>::GROUP > 1) {
        // variational systems: Q and P live in two lanes of the group; one decision per group (a non-finite
        // column must stop every thread of the group)
        constexpr int G = Sys<SYS>::GROUP;
        const unsigned base = group_base<G>(), mask = group_mask<G>();
        const double Q = __shfl_sync(mask, u[0], (int)base + Sys<SYS>::QI), P = __shfl_sync(mask, u[0], (int)base + Sys<SYS>::PI);
        const double r2 = __dsub_rn(__dsub_rn(4.0, __dmul_rn(P, P)), __dmul_rn(Q, Q));
        bool nonfinite = false;
#pragma unroll
        for (int i = 0; i < Sys<SYS>::NV; ++i) nonfinite |= ((__double2hiint(u[i]) & 0x7ff00000) == 0x7ff00000);
        const bool bad = nonfinite || !(r2 > 0.0);
        return (__ballot_sync(mask, bad) & mask) != 0u;
    } else if constexpr (SYS == DTWA_SYS_DICKE_BLOCH) {
        // no chart boundary; a state is bad when it is non-finite or sits on the pole jz >= 1 of the (Q,P) chart, where the
        // observables' coordinates do not exist (4 - Q^2 - P^2 = 2 (1 - jz) <= 0, the reference's out_of_bounds)
        bool nonfinite = false;
#pragma unroll
        for (int i = 0; i < Sys<SYS>::NV; ++i) nonfinite |= ((__double2hiint(u[i]) & 0x7ff00000) == 0x7ff00000);
        return nonfinite || !(u[2] < 1.0);
    } else {
        const double Q = u[Sys<SYS>::QI], P = u[Sys<SYS>::PI];
        const double r2 = __dsub_rn(__dsub_rn(4.0, __dmul_rn(P, P)), __dmul_rn(Q, Q));
        bool nonfinite = false;
#pragma unroll
        for (int i = 0; i < Sys<SYS>::NV; ++i)
            if (i != Sys<SYS>::QI && i != Sys<SYS>::PI) nonfinite |= ((__double2hiint(u[i]) & 0x7ff00000) == 0x7ff00000);
        return nonfinite || !(r2 > 0.0);
    }
}

// ------------------------------------------------------------------------------------------------
// fixed-step steppers
struct StepTab {  // one row per output interval (built on the host, dtwa_api.cu build_step_table)
    double h, hh, h3, h6;      // h, h/2, h/3, h/6
    double hw, hhw, h3w, h6w;  // the same times the oscillator frequency (Dicke: dq/dt = w p is folded into the RK4 weights)
    int32_t n;
    int32_t pad;
};

template <int SYS>
__device__ __forceinline__ void rk4_step(const SysPar &p, double (&u)[Sys<SYS>::NV], const StepTab &r)
{
    constexpr int NV = Sys<SYS>::NV;
    double k[NV], y[NV], acc[NV];
    if constexpr (SYS == DTWA_SYS_DICKE) {
        // component 1 (q): k = w * p_stage is never formed; its weights carry the factor w
        Sys<SYS>::template rhs<true>(p, u, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            y[i] = fma(i == 1 ? r.hhw : r.hh, k[i], u[i]);
            acc[i] = fma(i == 1 ? r.h6w : r.h6, k[i], u[i]);
        }
        Sys<SYS>::template rhs<true>(p, y, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            y[i] = fma(i == 1 ? r.hhw : r.hh, k[i], u[i]);
            acc[i] = fma(i == 1 ? r.h3w : r.h3, k[i], acc[i]);
        }
        Sys<SYS>::template rhs<true>(p, y, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            y[i] = fma(i == 1 ? r.hw : r.h, k[i], u[i]);
            acc[i] = fma(i == 1 ? r.h3w : r.h3, k[i], acc[i]);
        }
        Sys<SYS>::template rhs<true>(p, y, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) u[i] = fma(i == 1 ? r.h6w : r.h6, k[i], acc[i]);
    } else {
        Sys<SYS>::rhs(p, u, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            y[i] = fma(r.hh, k[i], u[i]);
            acc[i] = fma(r.h6, k[i], u[i]);
        }
        Sys<SYS>::rhs(p, y, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            y[i] = fma(r.hh, k[i], u[i]);
            acc[i] = fma(r.h3, k[i], acc[i]);
        }
        Sys<SYS>::rhs(p, y, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            y[i] = fma(r.h, k[i], u[i]);
            acc[i] = fma(r.h3, k[i], acc[i]);
        }
        Sys<SYS>::rhs(p, y, k);
#pragma unroll
        for (int i = 0; i < NV; ++i) u[i] = fma(r.h6, k[i], acc[i]);
    }
}

// coefficient tables of the embedded pairs, in constant memory (see tools/gen_tableaux.py)
#define DTWA_TABLEAU_CONSTANTS
#include "dp5_body.inc"
#include "tsit5_body.inc"
#include "dop853_body.inc"
#undef DTWA_TABLEAU_CONSTANTS

#define RHS(in, out) Sys<SYS>::rhs(p, in, out)
#define DTWA_STEP_H(i) h            // what the generated stage code is instantiated with: the step length,
#define DTWA_REAL double             // the arithmetic type (trial2 below: D2, two trajectories side by side)
#define DTWA_CMUL(c, x) ((c) * (x))  // and coefficient * value

// one DOP853 step with the 8th-order weights only (fixed step)
template <int SYS>
__device__ __forceinline__ void rk8_step(const SysPar &p, double (&y)[Sys<SYS>::NV], double h)
{
    constexpr int NV = Sys<SYS>::NV;
    double k0[NV];
    RHS(y, k0);
    // stages 1..11 and the weighted sum, generated straight-line code; the FSAL evaluation and
    // the error combinations it also emits are dead here and removed by the compiler
#include "dop853_body.inc"
    (void)e5;
    (void)e3;
    (void)dy;
#pragma unroll
    for (int i = 0; i < NV; ++i) y[i] = ynew[i];
#undef DOP853_KNEW
}

// ------------------------------------------------------------------------------------------------
// embedded adaptive pairs.  Error norm: RMS of err_i / (tol + max(|y_i|,|ynew_i|) tol) (OrdinaryDiffEq's
// calculate_residuals + ODE_DEFAULT_NORM and SciPy's rk.py agree on it), DOP853 with Hairer's 5/3 combination.
// Two step-size controllers (CtrlPar::kind):
//   DTWA_CTRL_ODE   the reference's: OrdinaryDiffEq v5 (solve() in src/ClassicalSystems.jl:138).  PI controller
//                   q = EEst^beta1 / qold^beta2 / gamma clamped to [1/qmax, 1/qmin]; accept when EEst <= 1, then
//                   dt <- dt/q and qold <- max(EEst, qoldinit); after a rejection dt <- dt / min(1/qmin, EEst^beta1/gamma);
//                   after an isoutofdomain rejection dt <- dt qmin.  Gains: the algorithm's defaults in alg_utils.jl
//                   (DP5 0.17/0.04, DP8 1/8 / 0 with qmin 1/3 qmax 6, order-8 generic = TsitPap8/Vern8 7/80 / 1/20) or the
//                   caller's (solve's beta1/beta2/qmin/qmax/gamma/qoldinit keywords).
//   DTWA_CTRL_SCIPY SciPy's RungeKutta._step_impl: factor = 0.9 err^(-1/(q+1)) in [0.2, 10], no growth right after a
//                   rejection, accept when err < 1 -- what tests/golden/scipy_pairs.json pins the tableaux with.
// Both under the reference's solver contract: dtmin with force_dtmin, isoutofdomain, maxiters.
// The powers only size the next step (never the accept/reject decision, taken on err in full double precision), so
// single precision through the MUFU units is ample.
// One branch-free rule covers both (lg2/ex2 through the MUFU units, no division, no libm slow paths: the controller is a
// serial, latency-bound tail of every attempt during which the warp feeds nothing to the FP64 pipe):
//   le = lg2(err);  accepted step: h *= min(fac_max, max(fac_min_acc, g 2^-(b1 le + nb2 lqold)))   [capped by post_reject_cap
//   right after a rejection];  rejected step: h *= max(fac_min_rej, g 2^-(b1 le));  accept when err < acc_thr.
//   ODE:   b1 = beta1, nb2 = -beta2, g = gamma, [fac_min_acc, fac_max] = [qmin, qmax], fac_min_rej = qmin, acc_thr = 1 + ulp,
//          lqold = lg2(max(EEst, qoldinit)) of the last accepted un-clipped step, post_reject_cap = qmax
//   SciPy: b1 = 1/(q+1), nb2 = 0, g = 0.9, fac_max = 10, fac_min_acc = 0, fac_min_rej = 0.2, acc_thr = 1, post_reject_cap = 1
struct CtrlPar {
    int32_t kind;
    float b1, nb2, g;
    float fac_max, fac_min_acc, fac_min_rej, post_reject_cap;
    float lqoldinit;
    float pad;
    double acc_thr;
    double dom_shrink;             // step factor after an isoutofdomain rejection: qmin (SciPy mode: 0.2)
};

__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// lg2 of a non-negative double from its bits: exponent + lg2 of the leading 24 bits of the significand (integer pipe + one
// MUFU.LG2; no conversion that could overflow or flush).  0 and denormals give about -1023, +Inf and NaN about 1024.
__device__ __forceinline__ float lg2_of_double(double x)
{
    const int hi = __double2hiint(x);
    const unsigned lo = (unsigned)__double2loint(x);
    const int e = ((hi >> 20) & 0x7ff) - 1023;
    const float m = __int_as_float(0x3f800000 | ((hi & 0xfffff) << 3) | (int)(lo >> 29));
    return (float)e + lg2_approx(m);
}
// lg2 of the error estimate for the step-size controller, clamped like err in [1e-30, 1e30]
#define DTWA_LE_CLAMP 99.65784f
__device__ __forceinline__ float clamp_le(float le) { return fminf(fmaxf(le, -DTWA_LE_CLAMP), DTWA_LE_CLAMP); }

template <int SYS, int ALG>
struct Adaptive {
    static constexpr int NV = Sys<SYS>::NV;
    static constexpr bool FIVE = (ALG == DTWA_ALG_DP5 || ALG == DTWA_ALG_TSIT5);   // the two 7-stage FSAL pairs of order 5(4)
    static constexpr int ERR_ORDER = FIVE ? 4 : 7;
    double t, habs;
    double hlast;      // length of the last accepted step (interpolated outputs)
    double k0[NV];
    float lqold;       // PI controller memory: lg2 of the error estimate of the last accepted (un-clipped) step
    bool rejected;

    __device__ __forceinline__ void init(const SysPar &p, const double (&y)[NV], double tend, double tol, const CtrlPar &c)
    {
        t = 0.0;
        rejected = false;
        lqold = c.lqoldinit;
        RHS(y, k0);
        // Hairer's starting step (scipy common.select_initial_step, direction +1, max_step inf)
        double sc[NV], a[NV];
        if (tend == 0.0) {
            habs = 0.0;
            return;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) sc[i] = fma(fabs(y[i]), tol, tol);
#pragma unroll
        for (int i = 0; i < NV; ++i) a[i] = y[i] / sc[i];
        const double d0 = group_rms<SYS>(a);
#pragma unroll
        for (int i = 0; i < NV; ++i) a[i] = k0[i] / sc[i];
        const double d1 = group_rms<SYS>(a);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, tend);
        double y1[NV], f1[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) y1[i] = fma(h0, k0[i], y[i]);
        RHS(y1, f1);
#pragma unroll
        for (int i = 0; i < NV; ++i) a[i] = (f1[i] - k0[i]) / sc[i];
        const double d2 = group_rms<SYS>(a) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15)
            h1 = fmax(1e-6, h0 * 1e-3);
        else
            h1 = pow(0.01 / fmax(d1, d2), 1.0 / (ERR_ORDER + 1));
        double h = fmin(100.0 * h0, h1);
        if (!(h == h)) h = 1e-6;
        habs = fmin(h, tend);
    }

    // The stages and the error estimate of ONE trial step of length h from (y, k0): no decision, no state change.  Shared by
    // attempt() below and by the software-pipelined hot loop of the ensemble kernels, so that both take bit-identical steps.
    // `le` = lg2(err) for the step-size controller, taken from the FACTORS of err in parallel with the square root that err itself
    // needs (lg2 err = lg2|h| + lg2 n5 - lg2(den)/2 for the 5/3 combination, lg2(mean square)/2 for the 7-stage pair), so that the
    // controller's chain does not wait for rsqrt -> product -> conversion.  MEASURED (call 23, -DDTWA_LE_FROM_FACTORS): +2.8 % for
    // DOP853 under TsitPap8's gains, -4.7 % under DP8's gains and for DP5, a 10^4-trajectory call 4.93 instead of 4.84 ms -- the
    // extra integer/MUFU instructions cost what the shorter chain returns.  attempt() therefore overrides it with lg2 of the
    // finished estimate (the compiler drops the unused factor form); the two-stream and pipelined experiments do the same.
    __device__ __forceinline__ void trial(const SysPar &p, const double (&y)[NV], const double h, const double tol, double (&yn)[NV],
                                          double (&kn)[NV], double &err, float &le) const
    {
        if constexpr (ALG == DTWA_ALG_DP5) {
#include "dp5_body.inc"
            (void)dy;
            double en[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double sc = fma(fmax(fabs(y[i]), fabs(ynew[i])), tol, tol);
                en[i] = e5[i] * h * rcp_fast(sc);
                yn[i] = ynew[i];
                kn[i] = DP5_KNEW[i];
            }
            const double msq = group_meansq<SYS>(en);
            err = sqrt_fast(msq);
            le = clamp_le(0.5f * lg2_of_double(msq));
#undef DP5_KNEW
        } else if constexpr (ALG == DTWA_ALG_TSIT5) {
#include "tsit5_body.inc"
            (void)dy;
            double en[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double sc = fma(fmax(fabs(y[i]), fabs(ynew[i])), tol, tol);
                en[i] = e5[i] * h * rcp_fast(sc);
                yn[i] = ynew[i];
                kn[i] = TSIT5_KNEW[i];
            }
            const double msq = group_meansq<SYS>(en);
            err = sqrt_fast(msq);
            le = clamp_le(0.5f * lg2_of_double(msq));
#undef TSIT5_KNEW
        } else {
#include "dop853_body.inc"
            double n5 = 0.0, n3 = 0.0;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double isc = rcp_fast(fma(fmax(fabs(y[i]), fabs(ynew[i])), tol, tol));  // one reciprocal for both estimators
                const double a5 = e5[i] * isc, a3 = e3[i] * isc;
                if (norm_counted<SYS>(i)) {
                    n5 = fma(a5, a5, n5);
                    n3 = fma(a3, a3, n3);
                }
                yn[i] = ynew[i];
                kn[i] = DOP853_KNEW[i];
            }
            n5 = group_sum<SYS>(n5);
            n3 = group_sum<SYS>(n3);
            {
                double is, sq;
                const double den = fma(0.01, n3, n5) * (double)Sys<SYS>::NORM_N;
                rsqrt_sqrt_pair(den, is, sq);
                const double e = fabs(h) * n5 * is;
                err = (n5 == 0.0 && n3 == 0.0) ? 0.0 : e;   // (0 * Inf otherwise)
                le = clamp_le(lg2_of_double(fabs(h)) + fmaf(-0.5f, lg2_of_double(den), lg2_of_double(n5)));
            }
#undef DOP853_KNEW
        }
    }

    // The same for TWO independent trajectories of one lane (a, b): the generated stage code and the right-hand side are
    // instantiated for the arithmetic type D2, so that every operation is issued for a and then for b and the two dependency
    // chains (stage sum -> rsqrt -> right-hand side, ~11 dependent FP64 instructions per stage) interleave instruction by
    // instruction in ONE basic block.  Every arithmetic operation is the one trial() executes for the same trajectory:
    // bit-identical steps.
    __device__ __forceinline__ static void trial2(const SysPar &p, const Adaptive &A0, const Adaptive &A1, const double (&ya)[NV], const double (&yb)[NV],
                                                  const double ha, const double hb, const double tol, double (&yna)[NV], double (&ynb)[NV],
                                                  double (&kna)[NV], double (&knb)[NV], double &erra, double &errb, float &lea, float &leb)
    {
        D2 y[NV], k0[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            y[i] = D2{ya[i], yb[i]};
            k0[i] = D2{A0.k0[i], A1.k0[i]};
        }
        const D2 h2{ha, hb};
#undef DTWA_STEP_H
#undef DTWA_REAL
#undef DTWA_CMUL
#define DTWA_STEP_H(i) h2
#define DTWA_REAL D2
#define DTWA_CMUL(c, x) dm(c, x)
        if constexpr (ALG == DTWA_ALG_DP5) {
#include "dp5_body.inc"
            (void)dy;
            double ena[NV], enb[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double sca = fma(fmax(fabs(y[i].a), fabs(ynew[i].a)), tol, tol), scb = fma(fmax(fabs(y[i].b), fabs(ynew[i].b)), tol, tol);
                ena[i] = e5[i].a * ha * rcp_fast(sca);
                enb[i] = e5[i].b * hb * rcp_fast(scb);
                yna[i] = ynew[i].a;
                ynb[i] = ynew[i].b;
                kna[i] = DP5_KNEW[i].a;
                knb[i] = DP5_KNEW[i].b;
            }
            const double msqa = group_meansq<SYS>(ena), msqb = group_meansq<SYS>(enb);
            erra = sqrt_fast(msqa);
            errb = sqrt_fast(msqb);
            lea = clamp_le(0.5f * lg2_of_double(msqa));
            leb = clamp_le(0.5f * lg2_of_double(msqb));
#undef DP5_KNEW
        } else if constexpr (ALG == DTWA_ALG_TSIT5) {
#include "tsit5_body.inc"
            (void)dy;
            double ena[NV], enb[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double sca = fma(fmax(fabs(y[i].a), fabs(ynew[i].a)), tol, tol), scb = fma(fmax(fabs(y[i].b), fabs(ynew[i].b)), tol, tol);
                ena[i] = e5[i].a * ha * rcp_fast(sca);
                enb[i] = e5[i].b * hb * rcp_fast(scb);
                yna[i] = ynew[i].a;
                ynb[i] = ynew[i].b;
                kna[i] = TSIT5_KNEW[i].a;
                knb[i] = TSIT5_KNEW[i].b;
            }
            const double msqa = group_meansq<SYS>(ena), msqb = group_meansq<SYS>(enb);
            erra = sqrt_fast(msqa);
            errb = sqrt_fast(msqb);
            lea = clamp_le(0.5f * lg2_of_double(msqa));
            leb = clamp_le(0.5f * lg2_of_double(msqb));
#undef TSIT5_KNEW
        } else {
#include "dop853_body.inc"
            double n5a = 0.0, n3a = 0.0, n5b = 0.0, n3b = 0.0;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double isca = rcp_fast(fma(fmax(fabs(y[i].a), fabs(ynew[i].a)), tol, tol));
                const double iscb = rcp_fast(fma(fmax(fabs(y[i].b), fabs(ynew[i].b)), tol, tol));
                const double a5a = e5[i].a * isca, a3a = e3[i].a * isca, a5b = e5[i].b * iscb, a3b = e3[i].b * iscb;
                if (norm_counted<SYS>(i)) {
                    n5a = fma(a5a, a5a, n5a);
                    n5b = fma(a5b, a5b, n5b);
                    n3a = fma(a3a, a3a, n3a);
                    n3b = fma(a3b, a3b, n3b);
                }
                yna[i] = ynew[i].a;
                ynb[i] = ynew[i].b;
                kna[i] = DOP853_KNEW[i].a;
                knb[i] = DOP853_KNEW[i].b;
            }
            n5a = group_sum<SYS>(n5a);
            n5b = group_sum<SYS>(n5b);
            n3a = group_sum<SYS>(n3a);
            n3b = group_sum<SYS>(n3b);
            D2 is, sq;
            const D2 den{fma(0.01, n3a, n5a) * (double)Sys<SYS>::NORM_N, fma(0.01, n3b, n5b) * (double)Sys<SYS>::NORM_N};
            rsqrt_sqrt_pair(den, is, sq);
            const double ea = fabs(ha) * n5a * is.a, eb = fabs(hb) * n5b * is.b;
            erra = (n5a == 0.0 && n3a == 0.0) ? 0.0 : ea;
            errb = (n5b == 0.0 && n3b == 0.0) ? 0.0 : eb;
            lea = clamp_le(lg2_of_double(fabs(ha)) + fmaf(-0.5f, lg2_of_double(den.a), lg2_of_double(n5a)));
            leb = clamp_le(lg2_of_double(fabs(hb)) + fmaf(-0.5f, lg2_of_double(den.b), lg2_of_double(n5b)));
#undef DOP853_KNEW
        }
#undef DTWA_STEP_H
#undef DTWA_REAL
#undef DTWA_CMUL
#define DTWA_STEP_H(i) h
#define DTWA_REAL double
#define DTWA_CMUL(c, x) ((c) * (x))
    }

    // one attempted step towards tstop.  Returns 0 = keep going, 1 = trajectory failed (DOMAIN).
    // `accepted` reports whether y/t advanced.  KEEP: an accepted step also hands back the state and the derivative at its
    // START (yprev, fprev) and its length (hlast) -- what the Hermite interpolant of interpolated outputs needs; they take
    // the registers of the trial values, which are dead by then.
    template <bool KEEP = false>
    __device__ __forceinline__ int attempt(const SysPar &p, double (&y)[NV], double tstop, double tol, double dtmin,
                                           bool &accepted, const CtrlPar &c, double *yprev = nullptr, double *fprev = nullptr)
    {
        accepted = false;
        bool forced = false;
        if (habs < dtmin) {
            habs = dtmin;
            forced = true;
        }
        double h = habs;
        bool clipped = false;
        const bool last = (t + h >= tstop);
        if (last) {
            clipped = (tstop - t) < habs;
            h = tstop - t;
        }
        double err;
        float le;
        double yn[NV], kn[NV];
        trial(p, y, h, tol, yn, kn, err, le);
#ifndef DTWA_LE_FROM_FACTORS   // default: lg2 of the finished error estimate (trial()'s factor form is an opt-in experiment)
        le = lg2_approx(fminf(fmaxf((float)err, 1e-30f), 1e30f));
#endif
        const bool bad = state_bad<SYS>(yn) || !(err == err) || isinf(err);
        if (bad) {
            if (forced) return 1;
            habs = h * c.dom_shrink;
            rejected = true;
            return 0;
        }
        const float e_rej = ex2_approx(-c.b1 * le);
        const float e_acc = ex2_approx(-fmaf(c.b1, le, c.nb2 * lqold));
        float fac_acc = fminf(c.fac_max, fmaxf(c.fac_min_acc, c.g * e_acc));   // err == 0: e_acc = +inf, clamped to fac_max
        if (rejected) fac_acc = fminf(fac_acc, c.post_reject_cap);
        const float fac_rej = fmaxf(c.fac_min_rej, c.g * e_rej);
        const bool acc = forced || err < c.acc_thr;
        double hn = h * (double)(acc ? fac_acc : fac_rej);
        if (acc) {
            if (clipped) {
                // a step cut short to land on ts[k] (the reference interpolates there, src/TWA.jl:180) does not feed the
                // controller: the un-clipped proposal stands unless the new one is larger
                if (hn < habs) hn = habs;
            } else
                lqold = fmaxf(le, c.lqoldinit);
            t = last ? tstop : t + h;
            if constexpr (KEEP) {
                hlast = h;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    yprev[i] = y[i];
                    fprev[i] = k0[i];
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                y[i] = yn[i];
                k0[i] = kn[i];
            }
        }
        habs = hn;
        rejected = !acc;
        accepted = acc;
        return 0;
    }
};
#undef RHS
#undef DTWA_STEP_H
#undef DTWA_REAL
#undef DTWA_CMUL

// 3rd-order Hermite interpolant over an accepted step: OrdinaryDiffEq's hermite_interpolant, what the reference's
// FunctionCallingCallback(funcat = ts) evaluates at the output times (src/TWA.jl:180) for algorithms without a dense output
// of their own (TsitPap8, the reference default):
//   u(Theta) = (1 - Theta) y0 + Theta y1 + Theta (Theta - 1) ((1 - 2 Theta)(y1 - y0) + (Theta - 1) h f0 + Theta h f1)
template <int NV>
__device__ __forceinline__ void hermite_interpolant(const double *y0, const double *f0, const double (&y1)[NV], const double (&f1)[NV], double h,
                                                    double theta, double (&u)[NV])
{
    const double tm1 = theta - 1.0, a = theta * tm1, b = 1.0 - 2.0 * theta;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double dy = y1[i] - y0[i];
        const double inner = fma(b, dy, h * fma(tm1, f0[i], theta * f1[i]));
        u[i] = fma(a, inner, fma(theta, dy, y0[i]));
    }
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  key = seed, counter = (idx, attempt, block).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    const uint64_t m = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return __dmul_rn(__dadd_rn((double)m, 0.5), 1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void uniforms4(uint64_t seed, uint64_t idx, uint32_t attempt, double (&U)[4])
{
    uint32_t r[4];
    philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), attempt, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    U[0] = u53(r[0], r[1]);
    U[1] = u53(r[2], r[3]);
    philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), attempt, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    U[2] = u53(r[0], r[1]);
    U[3] = u53(r[2], r[3]);
}

struct SamplerPar {
    int32_t kind;
    int32_t nv;
    double center[4];
    double hbar;
    uint64_t seed;
    uint64_t offset;
    const double *u0;  // device pointer for array samplers
};

#define DTWA_TWO_PI 6.283185307179586476925286766559
#define DTWA_PI 3.14159265358979323846

__device__ __forceinline__ double mod2pi_pos(double x)
{
    double r = fmod(x, DTWA_TWO_PI);
    if (r < 0.0) r += DTWA_TWO_PI;
    return r;
}
__device__ __forceinline__ double theta_of_QP(double Q, double P) { return acos(1.0 - (Q * Q + P * P) / 2.0); }
__device__ __forceinline__ double phi_of_QP(double Q, double P) { return mod2pi_pos(atan2(-P, Q)); }

// src/TWA.jl:708-745 sample(): Theta ~ Rayleigh(sigma) by inversion of U1, azimuth 2 pi U2, rotation
// R(theta,phi,theta0,phi0) of :712-714, then (Q,P) of src/PhaseSpaces.jl:16,34 -- formula for formula.
__device__ __noinline__ void sample_su2(double Q0, double P0, double hbar, double U1, double U2, double &Q, double &P)
{
    const double sigma = sqrt(hbar / 2.0);
    const double th0 = theta_of_QP(Q0, P0), ph0 = phi_of_QP(Q0, P0);
    const double th = sigma * sqrt(-2.0 * log(U1));
    const double ph = U2 * 2.0 * DTWA_PI;
    const double sth = sin(th), cth = cos(th), sph = sin(ph), cph = cos(ph);
    const double st0 = sin(th0), ct0 = cos(th0), sp0 = sin(ph0), cp0 = cos(ph0);
    const double a = ct0 * cph * sth - cth * st0;
    const double thr = atan2(sqrt(a * a + sth * sth * sph * sph), cth * ct0 + cph * sth * st0);
    const double phr = mod2pi_pos(atan2(-cp0 * sth * sph - ct0 * cph * sth * sp0 + cth * st0 * sp0,
                                        -ct0 * cph * cp0 * sth + cth * cp0 * st0 + sth * sph * sp0));
    const double rad = sqrt(2.0 * (1.0 - cos(thr)));
    Q = rad * cos(phr);
    P = -rad * sin(phr);
}

// src/TWA.jl:689-700: two independent normals (Box-Muller on U3, U4)
__device__ __noinline__ void sample_hw(double q0, double p0, double hbar, double U3, double U4, double &q, double &p)
{
    const double sigma = sqrt(hbar / 2.0);
    const double r = sigma * sqrt(-2.0 * log(U3));
    const double a = U4 * 2.0 * DTWA_PI;
    q = q0 + r * cos(a);
    p = p0 + r * sin(a);
}

template <int NV>
__device__ __forceinline__ void draw_initial(const SamplerPar &sp, uint64_t local_idx, uint32_t attempt, double (&u)[NV])
{
    if (sp.kind == DTWA_SAMPLER_HOST_ARRAY || sp.kind == DTWA_SAMPLER_DEVICE_ARRAY) {
        const double *src = sp.u0 + local_idx * NV;
        if constexpr (NV == 4) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(src));
            const double2 b = __ldg(reinterpret_cast<const double2 *>(src) + 1);
            u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y;
        } else {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(src));
            u[0] = a.x; u[1] = a.y;
        }
        return;
    }
    double U[4];
    uniforms4(sp.seed, sp.offset + local_idx, attempt, U);
    if constexpr (NV == 4) {
        sample_su2(sp.center[0], sp.center[2], sp.hbar, U[0], U[1], u[0], u[2]);
        sample_hw(sp.center[1], sp.center[3], sp.hbar, U[2], U[3], u[1], u[3]);
    } else {
        if (sp.kind == DTWA_SAMPLER_COHERENT_HW)
            sample_hw(sp.center[0], sp.center[1], sp.hbar, U[2], U[3], u[0], u[1]);
        else
            sample_su2(sp.center[0], sp.center[1], sp.hbar, U[0], U[1], u[0], u[1]);
    }
}

// probability densities: src/TWA.jl:692-694 (MvNormal pdf), :719-736 (W_theta; 0.0 where the
// reference's try/catch swallows an acos DomainError), :757 (product)
__device__ __noinline__ double pdf_hw(double q0, double p0, double hbar, double q, double p)
{
    const double s2 = hbar / 2.0;
    return exp(-((q - q0) * (q - q0) + (p - p0) * (p - p0)) / (2.0 * s2)) / (2.0 * DTWA_PI * s2);
}
__device__ __noinline__ double pdf_su2(double Q0, double P0, double hbar, double Q, double P)
{
    // src/TWA.jl:719-736: W = N(Theta; 0, sigma) sqrt(2 pi) sigma / (2 pi sigma^2) with Theta the angle between the
    // Bloch vectors of (Q,P) and of the centre, cos(Theta) = cos(th) cos(th0) + sin(th) sin(th0) cos(ph - ph0).
    // With cos(th) = 1 - R/2, sin(th) (cos ph, sin ph) = sqrt(1 - R/4) (Q, -P), R = Q^2 + P^2 (src/PhaseSpaces.jl:49,63)
    // that is  a a0 + (Q Q0 + P P0) sqrt((1 - R/4)(1 - R0/4)) : one sqrt instead of two acos, two atan2 and six
    // sin/cos -- this runs for every trajectory at every output time of survival_probability.  0.0 where the
    // reference's try/catch swallows an acos DomainError.
    const double sigma = sqrt(hbar / 2.0);
    const double R = Q * Q + P * P, R0 = Q0 * Q0 + P0 * P0;
    const double a = 1.0 - R / 2.0, a0 = 1.0 - R0 / 2.0;
    if (!(a >= -1.0 && a <= 1.0) || !(a0 >= -1.0 && a0 <= 1.0)) return 0.0;
    const double c = a * a0 + (Q * Q0 + P * P0) * sqrt((1.0 - R / 4.0) * (1.0 - R0 / 4.0));
    if (!(c >= -1.0 && c <= 1.0)) return 0.0;
    const double T = acos(c);
    const double g = exp(-(T * T) / (2.0 * sigma * sigma)) / (sqrt(2.0 * DTWA_PI) * sigma);
    return g * sqrt(2.0 * DTWA_PI) * sigma * (1.0 / (2.0 * DTWA_PI * sigma * sigma));
}
template <int NV>
__device__ __forceinline__ double pdf_eval(const SamplerPar &sp, const double (&u)[NV])
{
    if constexpr (NV == 4) {
        return pdf_su2(sp.center[0], sp.center[2], sp.hbar, u[0], u[2]) *
               pdf_hw(sp.center[1], sp.center[3], sp.hbar, u[1], u[3]);
    } else {
        if (sp.kind == DTWA_SAMPLER_COHERENT_HW) return pdf_hw(sp.center[0], sp.center[1], sp.hbar, u[0], u[1]);
        return pdf_su2(sp.center[0], sp.center[1], sp.hbar, u[0], u[1]);
    }
}

// ------------------------------------------------------------------------------------------------
// src/TWA.jl:53-64 scale_to_index, un-fused, round-half-even; 1-based, 0 = dropped
struct RangeD {
    double start, stop, lenm1;
    int64_t len;
    double c;     // lenm1 / (stop - start), for the fast path
    double pad;
};
// The reference's formula is Int(round((v - start)/(stop - start) * (len - 1) + 1)): a correctly rounded division, a
// multiplication and an addition, then round-half-even.  A division costs ~12 FP64-pipe instructions and runs for every
// sample of every histogram, so the index is first computed as rint(fma(v - start, c, 1)) with c = (len-1)/(stop-start):
// both values lie within a few ulp of the exact real number, hence they round to the same integer unless that number sits
// within 16 ulp of a rounding boundary k + 1/2 -- only then (probability ~1e-13 len per sample) the reference's own
// operation sequence decides.  Bit-exact with the un-fused formula by construction (tests: edges, ties, NaN, +-Inf).
__device__ __forceinline__ int scale_to_index(double v, const RangeD &r)
{
    if (!(v >= r.start && v <= r.stop)) return 0;   // outside, or NaN
    if (r.len == 1) return 1;
    const double d = __dsub_rn(v, r.start);
    const double y = fma(d, r.c, 1.0);
    const double k = rint(y);
    if (0.5 - fabs(y - k) > y * 0x1p-48) return (int)k;
    const double x = __ddiv_rn(d, __dsub_rn(r.stop, r.start));
    const double z = __dadd_rn(__dmul_rn(x, r.lenm1), 1.0);
    if (!(z == z)) return 0;   // degenerate range (start == stop, len > 1): 0/0 -- the reference would throw; dropped
    return (int)rint(z);
}

// ------------------------------------------------------------------------------------------------
// Observable VM: an accumulator machine with a small operand stack, one program shared by all threads
// of the launch (uniform control flow).  Every arithmetic op is a single correctly rounded IEEE
// operation (no contraction), so + - * / sqrt and integer powers agree bit for bit with the CPU
// oracle's evaluator.  Operand forms: stack (pop), constant table entry b, state variable a.
enum VmOp : uint8_t {
    OP_END = 0,
    OP_LDV, OP_LDC, OP_LDPDF,      // acc = var a | const b | pdf of the sampling distribution at the current state
    OP_PLDV, OP_PLDC, OP_PLDPDF,   // push acc first
    OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_ATAN2, OP_MIN, OP_MAX,   // acc = pop() (op) acc
    OP_ADDC, OP_SUBC, OP_RSUBC, OP_MULC, OP_DIVC, OP_RDIVC,             // acc (op) const b ; R*: const (op) acc
    OP_ADDV, OP_SUBV, OP_RSUBV, OP_MULV, OP_DIVV, OP_RDIVV,             // acc (op) var a
    OP_NEG, OP_SQR, OP_CUBE, OP_IPOW, OP_SQRT, OP_SIN, OP_COS, OP_TAN, OP_EXP, OP_LOG, OP_ABS, OP_ACOS, OP_ASIN, OP_ATAN,
    // sinks (consume acc)
    OP_SUM,     // -> sums slot b
    OP_BIN,     // -> index register a (0/1) through range b
    OP_HIST,    // histogram b from the index register(s)
    OP_STORE    // -> out (dtwa_eval_expr)
};
struct VmIns {
    uint8_t op;
    uint8_t a;
    uint16_t b;
};
constexpr int VM_REG_STACK = 4;    // operand stack entries held in registers
constexpr int VM_STACK = 16;       // total operand stack depth (the rest overflows to local memory)
constexpr int VM_MAX_INS = 1024;
constexpr int VM_MAX_CONST = 256;
constexpr int VM_MAX_RANGE = 32;
constexpr int VM_MAX_HIST = 16;
constexpr int VM_MAX_SLOTS = 64;

__device__ __forceinline__ double ipow_rn(double x, int n)
{
    // left-to-right binary powering: x^2 = x*x, x^3 = (x*x)*x, x^4 = (x*x)*(x*x) ...
    if (n == 0) return 1.0;
    const bool neg = n < 0;
    unsigned m = neg ? (unsigned)(-n) : (unsigned)n;
    int top = 31 - __clz(m);
    double r = x;
    for (int bit = top - 1; bit >= 0; --bit) {
        r = __dmul_rn(r, r);
        if ((m >> bit) & 1u) r = __dmul_rn(r, x);
    }
    return neg ? __ddiv_rn(1.0, r) : r;
}

}  // namespace dtwa
