/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the TWA-ensemble hot path of saulpila/DickeModel.jl.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object; the product path (libdicketwa.so) never does and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference ships no golden vectors (test/runtests.jl:4-6 is an empty
 * testset), Julia is not installed in this image so the reference cannot be run, and its ODE
 * arithmetic lives in un-vendored packages (OrdinaryDiffEq "5", DiffEqCallbacks "2",
 * Distributions "0.23, 0.25"; Project.toml:29-46).  What pins this oracle instead (see
 * oracle/make_golden.py, tests/golden/): 40-digit mpmath Taylor truth trajectories, SciPy
 * DOP853/RK45 runs of the same tableaux, energy conservation, time reversal and sampler moments; and,
 * from the reference's own closed forms, the normal frequencies of src/ClassicalDicke.jl:139-147 (eigenvalues of
 * the fundamental matrix at minimum_energy_point) and the shell volume of :110-124 (sum of the quadrature).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off so that no FMA contraction happens and the
 * arithmetic is the reference's operation order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "tableaux_gen.h"

#define SYS_DICKE 0
#define SYS_LMG 1

#define ALG_RK4 0
#define ALG_RK8 1 /* fixed step with the DOP853 8th-order weights */
#define ALG_DP5 2
#define ALG_DOP853 3
#define ALG_TSIT5 4 /* Tsitouras 5(4): the same structure as DP5 (6 stages + FSAL, one 4th-order error estimator), its own tableau */
#define IS_FIVE(alg) ((alg) == ALG_DP5 || (alg) == ALG_TSIT5)
#define FIVE_A(alg) ((alg) == ALG_TSIT5 ? &TSIT5_A[0][0] : &DP5_A[0][0])
#define FIVE_B(alg) ((alg) == ALG_TSIT5 ? TSIT5_B : DP5_B)
#define FIVE_E(alg) ((alg) == ALG_TSIT5 ? TSIT5_E : DP5_E)

#define ST_OK 0
#define ST_DOMAIN 1   /* DomainError / NaN: the reference throws (src/ClassicalDicke.jl:12-15 sqrt of a negative) */
#define ST_MAXITERS 2 /* solver hit maxiters: no exception in the reference, later times get nothing */
#define ST_BAD_IC 3   /* src/ClassicalSystems.jl:100-102 */

#define SAMPLER_HOST 0
#define SAMPLER_HW 1
#define SAMPLER_SU2 2
#define SAMPLER_HWXSU2 3

static const double TWO_PI = 6.283185307179586476925286766559;

int orc_nvar(int kind) { return kind == SYS_DICKE ? 4 : 2; }

/* ------------------------------------------------------------------------------------------ */
/* a1/a3: equations of motion.  src/ClassicalDicke.jl:11-17 (dicke_step), src/ClassicalLMG.jl:8-11
 * (LMG_step).  par = [w0, w, g] (Dicke, order of ClassicalDickeSystem :50) or [Omega, xi] (LMG :34);
 * sigma = +-1 is the reference's `integate_backwards` parameter.  Returns 1 when the Dicke square
 * root would throw a DomainError (argument < 0; we also flag == 0, the out_of_bounds rule :18-20). */
int orc_rhs(int kind, const double *par, double sigma, const double *u, double *du)
{
    if (kind == SYS_DICKE) {
        double w0 = par[0], w = par[1], g = par[2];
        double Q = u[0], q = u[1], P = u[2], p = u[3];
        double r2 = 4 - P * P - Q * Q;
        int bad = !(r2 > 0);
        double s = sqrt(r2);
        du[0] = sigma * (P * w0 - g * P * q * Q / s);
        du[1] = sigma * p * w;
        du[2] = sigma * (g * q * (Q * Q) / s - g * q * s - Q * w0);
        du[3] = sigma * (-g * Q * s - q * w);
        return bad;
    } else {
        double Om = par[0], xi = par[1];
        double Q = u[0], P = u[1];
        du[0] = sigma * P * (Om - xi * (Q * Q) / 2);
        du[1] = sigma * (Q * (xi * (P * P) / 2 - (2 * xi + Om)) + xi * (Q * Q * Q));
        return 0;
    }
}

/* a2: src/ClassicalDicke.jl:18-20, src/ClassicalLMG.jl:12-14 */
int orc_out_of_bounds(int kind, const double *u)
{
    double Q = u[0], P = (kind == SYS_DICKE) ? u[2] : u[1];
    return 4 - P * P - Q * Q <= 0;
}

/* a5: src/ClassicalDicke.jl:80-88, src/ClassicalLMG.jl:63-70 */
double orc_hamiltonian(int kind, const double *par, const double *u)
{
    if (kind == SYS_DICKE) {
        double w0 = par[0], w = par[1], g = par[2];
        double Q = u[0], q = u[1], P = u[2], p = u[3];
        double r = 1 - (Q * Q + P * P) / 4;
        return (w0 / 2) * (Q * Q + P * P) + (w / 2) * (q * q + p * p) + 2 * g * q * Q * sqrt(r) - w0;
    } else {
        double Om = par[0], xi = par[1];
        double Q = u[0], P = u[1];
        return Om * (Q * Q + P * P) / 2 + xi * (Q * Q - (P * P) * (Q * Q) / 4 - (Q * Q) * (Q * Q) / 4) - Om; /* x^4 = (x*x)*(x*x), DESIGN.md */
    }
}

void orc_hamiltonian_batch(int kind, const double *par, const double *u, long n, double *h)
{
    int nv = orc_nvar(kind);
    for (long i = 0; i < n; ++i) h[i] = orc_hamiltonian(kind, par, u + i * nv);
}

static int nonfinite(const double *u, int nv)
{
    for (int i = 0; i < nv; ++i)
        if (!isfinite(u[i])) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Fixed-step table: the library steps exactly onto every ts[k] (SURVEY.md section 7 hard part 7;
 * the reference interpolates instead, src/TWA.jl:180).  Interval k runs from ts[k-1] (0 for k=0)
 * to ts[k] in nsub[k] equal sub-steps hsub[k] <= dt.  Returns -1 if ts is not sorted / negative. */
int orc_step_table(const double *ts, int nt, double dt, int *nsub, double *hsub)
{
    double tprev = 0.0;
    for (int k = 0; k < nt; ++k) {
        double d = ts[k] - tprev;
        if (!(d >= 0)) return -1;
        if (d == 0) {
            nsub[k] = 0;
            hsub[k] = 0;
        } else {
            double r = d / dt;
            long n = (long)ceil(r - 1e-9);
            if (n < 1) n = 1;
            nsub[k] = (int)n;
            hsub[k] = d / (double)n;
        }
        tprev = ts[k];
    }
    /* uniform grids (Julia ranges): one common sub-step, the library's detect_uniform rule */
    if (nt >= 2 && nsub[1] >= 1) {
        int n = nsub[1], ok = 1;
        double span = ts[nt - 1] - ts[0], mean = span / (double)(nt - 1);
        double noise = 8.0 * 2.220446049250313e-16 * fabs(ts[nt - 1]);
        for (int k = 1; k < nt && ok; ++k)
            if (nsub[k] != n || fabs((ts[k] - ts[k - 1]) - mean) > noise) ok = 0;
        if (ok && !(nsub[0] == 0 || (nsub[0] == n && fabs(ts[0] - mean) <= noise))) ok = 0;
        if (ok)
            for (int k = 0; k < nt; ++k)
                if (nsub[k]) hsub[k] = mean / (double)n;
    }
    return 0;
}

/* classical RK4, one step */
static int rk4_step(int kind, const double *par, double sg, double h, double *u)
{
    int nv = orc_nvar(kind), bad = 0;
    double k1[4], k2[4], k3[4], k4[4], y[4];
    bad |= orc_rhs(kind, par, sg, u, k1);
    for (int i = 0; i < nv; ++i) y[i] = u[i] + (h / 2) * k1[i];
    bad |= orc_rhs(kind, par, sg, y, k2);
    for (int i = 0; i < nv; ++i) y[i] = u[i] + (h / 2) * k2[i];
    bad |= orc_rhs(kind, par, sg, y, k3);
    for (int i = 0; i < nv; ++i) y[i] = u[i] + h * k3[i];
    bad |= orc_rhs(kind, par, sg, y, k4);
    for (int i = 0; i < nv; ++i) u[i] = u[i] + (h / 6) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    return bad;
}

/* generic explicit RK trial step from a dense tableau (scipy rk_step).  K has S+1 rows;
 * K[0] must hold f(y).  Computes ynew and K[S]=f(ynew). */
static int rk_trial(int kind, const double *par, double sg, int S, const double *A, int lda,
                    const double *B, double h, const double *y, double K[][4], double *ynew)
{
    int nv = orc_nvar(kind), bad = 0;
    double yt[4];
    for (int s = 1; s < S; ++s) {
        for (int i = 0; i < nv; ++i) {
            double dy = 0;
            for (int j = 0; j < s; ++j) dy += A[s * lda + j] * K[j][i];
            yt[i] = y[i] + dy * h;
        }
        bad |= orc_rhs(kind, par, sg, yt, K[s]);
    }
    for (int i = 0; i < nv; ++i) {
        double dy = 0;
        for (int j = 0; j < S; ++j) dy += B[j] * K[j][i];
        ynew[i] = y[i] + h * dy;
    }
    bad |= orc_rhs(kind, par, sg, ynew, K[S]);
    return bad;
}

static int rk8_fixed_step(int kind, const double *par, double sg, double h, double *u)
{
    int nv = orc_nvar(kind);
    double K[13][4], yn[4];
    int bad = orc_rhs(kind, par, sg, u, K[0]);
    bad |= rk_trial(kind, par, sg, 12, &DOP853_A[0][0], 12, DOP853_B, h, u, K, yn);
    for (int i = 0; i < nv; ++i) u[i] = yn[i];
    return bad;
}

static double rms(const double *x, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s) / sqrt((double)n);
}

/* Hairer's starting step as in scipy common.select_initial_step (direction +1, max_step inf) */
static double initial_step(int kind, const double *par, double sg, const double *y0, const double *f0,
                           double interval, int err_order, double tol)
{
    int nv = orc_nvar(kind);
    double sc[4], a[4], y1[4], f1[4];
    if (interval == 0) return 0;
    for (int i = 0; i < nv; ++i) sc[i] = tol + fabs(y0[i]) * tol;
    for (int i = 0; i < nv; ++i) a[i] = y0[i] / sc[i];
    double d0 = rms(a, nv);
    for (int i = 0; i < nv; ++i) a[i] = f0[i] / sc[i];
    double d1 = rms(a, nv);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > interval) h0 = interval;
    for (int i = 0; i < nv; ++i) y1[i] = y0[i] + h0 * f0[i];
    orc_rhs(kind, par, sg, y1, f1);
    for (int i = 0; i < nv; ++i) a[i] = (f1[i] - f0[i]) / sc[i];
    double d2 = rms(a, nv) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15)
        h1 = fmax(1e-6, h0 * 1e-3);
    else
        h1 = pow(0.01 / fmax(d1, d2), 1.0 / (err_order + 1));
    double h = fmin(100 * h0, h1);
    if (!(h == h)) h = 1e-6; /* NaN guard */
    return fmin(h, interval);
}

/* error norm of an embedded pair (scipy rk.py RungeKutta._estimate_error_norm / DOP853; OrdinaryDiffEq's calculate_residuals agrees) */
static double orc_error_norm(int alg, int nv, int S, double tol, double h, const double *u, const double *yn, double K[][4])
{
    double sc[4];
    for (int i = 0; i < nv; ++i) sc[i] = tol + fmax(fabs(u[i]), fabs(yn[i])) * tol;
    if (IS_FIVE(alg)) {
        double e[4];
        for (int i = 0; i < nv; ++i) {
            double s = 0;
            for (int j = 0; j <= S; ++j) s += FIVE_E(alg)[j] * K[j][i];
            e[i] = s * h / sc[i];
        }
        return rms(e, nv);
    }
    double n5 = 0, n3 = 0;
    for (int i = 0; i < nv; ++i) {
        double s5 = 0, s3 = 0;
        for (int j = 0; j <= S; ++j) {
            s5 += DOP853_E5[j] * K[j][i];
            s3 += DOP853_E3[j] * K[j][i];
        }
        s5 /= sc[i];
        s3 /= sc[i];
        n5 += s5 * s5;
        n3 += s3 * s3;
    }
    if (n5 == 0 && n3 == 0) return 0;
    return fabs(h) * n5 / sqrt((n5 + 0.01 * n3) * nv);
}

/* ------------------------------------------------------------------------------------------ */
/* a7 + the per-time callback of a20.  One trajectory, stepping exactly onto every ts[k]; calls
 * cb(user, k, u) at each output.  src/ClassicalSystems.jl:78-140 contract: abstol=reltol=tol,
 * dtmin (=tol by default) with force_dtmin, isoutofdomain step rejection, maxiters.  The adaptive
 * controller is SciPy's (rk.py _step_impl), not OrdinaryDiffEq's PI controller (not in the image).
 * Returns status; *fail_index = output index at which the trajectory stopped (nt when complete). */
typedef void (*orc_out_cb)(void *user, int k, const double *u);

typedef struct {
    int kind;
    double par[4];
    int backwards;
    int alg;
    double dt, tol, dtmin;
    long maxiters;
    /* step-size controller of the adaptive pairs.  0 = OrdinaryDiffEq's (the reference's solve(), src/ClassicalSystems.jl:138):
     * PI controller q = EEst^beta1 / qold^beta2 / gamma clamped to [1/qmax, 1/qmin], accept when EEst <= 1, qold = max(EEst,
     * qoldinit) after an accepted step, dt/min(1/qmin, EEst^beta1/gamma) after a rejected one, dt*qmin after an
     * isoutofdomain rejection (OrdinaryDiffEq v5 integrators/controllers + integrator_utils loopheader!/loopfooter!).
     * 1 = SciPy's (rk.py _step_impl): what tests/golden/scipy_pairs.json pins. */
    int controller;
    double beta1, beta2, qmin, qmax, gamma, qoldinit;
    /* 0: adaptive steps are clipped so that they land on every ts[k] (the library's default).  1: the reference's semantics
     * (FunctionCallingCallback(funcat=ts), src/TWA.jl:180): steps ignore the output times and the state at ts[k] is
     * OrdinaryDiffEq's 3rd-order Hermite interpolant over the accepted step that contains it; a step is clipped only at the
     * maxout-th output time ahead (maxout = min(32, half of a ring of min(256, nt) outputs): the library's ring rule) and at
     * the last one. */
    int interp;
} orc_problem;

#define CTRL_ODE 0
#define CTRL_SCIPY 1

int orc_solve_one(const orc_problem *pb, const double *u0, const double *ts, int nt, const int *nsub,
                  const double *hsub, orc_out_cb cb, void *user, int *fail_index, long *nacc,
                  long *nrej)
{
    int kind = pb->kind, nv = orc_nvar(kind);
    double sg = pb->backwards ? -1.0 : 1.0;
    const double *par = pb->par;
    double u[4];
    long acc = 0, rej = 0;
    int status = ST_OK, k = 0;
    memcpy(u, u0, sizeof(double) * nv);
    if (nonfinite(u, nv) || orc_out_of_bounds(kind, u)) {
        *fail_index = 0;
        *nacc = 0;
        *nrej = 0;
        return ST_BAD_IC;
    }
    if (pb->alg == ALG_RK4 || pb->alg == ALG_RK8) {
        for (k = 0; k < nt; ++k) {
            int bad = 0;
            for (int s = 0; s < nsub[k]; ++s) {
                bad |= (pb->alg == ALG_RK4) ? rk4_step(kind, par, sg, hsub[k], u)
                                            : rk8_fixed_step(kind, par, sg, hsub[k], u);
                acc++;
            }
            if (bad || nonfinite(u, nv) || orc_out_of_bounds(kind, u)) {
                status = ST_DOMAIN;
                break;
            }
            cb(user, k, u);
        }
    } else {
        int S = (IS_FIVE(pb->alg)) ? 6 : 12;
        int err_order = (IS_FIVE(pb->alg)) ? 4 : 7;
        double expo = -1.0 / (err_order + 1);
        double tol = pb->tol, dtmin = pb->dtmin;
        double K[13][4], yn[4];
        double t = 0, tend = ts[nt - 1];
        long attempts = 0;
        int have_f = 0;
        double h_abs = 0;
        int rejected = 0;
        double qold = pb->qoldinit, q11 = 1.0;
        if (pb->interp) {
            int wfull = nt < 256 ? nt : 256;
            int maxout = wfull >= 2 ? wfull / 2 : 1;
            if (maxout > 32) maxout = 32;
            double yprev[4], fprev[4], hlast = 0;
            k = 0;
            while (k < nt && status == ST_OK) {
                /* outputs due at the current time (before the first step: those at t = 0) */
                while (k < nt && !(t < ts[k])) {
                    if (nonfinite(u, nv) || orc_out_of_bounds(kind, u)) {
                        status = ST_DOMAIN;
                        break;
                    }
                    cb(user, k, u);
                    ++k;
                }
                if (k >= nt || status != ST_OK) break;
                if (!have_f) {
                    if (orc_rhs(kind, par, sg, u, K[0])) {
                        status = ST_DOMAIN;
                        break;
                    }
                    h_abs = initial_step(kind, par, sg, u, K[0], tend - t, err_order, tol);
                    have_f = 1;
                }
                if (attempts >= pb->maxiters) {
                    status = ST_MAXITERS;
                    break;
                }
                attempts++;
                int klim = k + maxout - 1 < nt - 1 ? k + maxout - 1 : nt - 1;
                double tlim = ts[klim];
                int forced = 0;
                if (h_abs < dtmin) {
                    h_abs = dtmin;
                    forced = 1;
                }
                double h = h_abs;
                int clipped = 0, last = 0;
                if (t + h >= tlim) {
                    clipped = (tlim - t) < h_abs;
                    h = tlim - t;
                    last = 1;
                }
                int bad;
                if (IS_FIVE(pb->alg))
                    bad = rk_trial(kind, par, sg, 6, FIVE_A(pb->alg), 6, FIVE_B(pb->alg), h, u, K, yn);
                else
                    bad = rk_trial(kind, par, sg, 12, &DOP853_A[0][0], 12, DOP853_B, h, u, K, yn);
                double err = orc_error_norm(pb->alg, nv, S, tol, h, u, yn, K);
                bad = bad || nonfinite(yn, nv) || !isfinite(err) || orc_out_of_bounds(kind, yn);
                if (bad) {
                    if (forced) {
                        status = ST_DOMAIN;
                        break;
                    }
                    h_abs = h * (pb->controller == CTRL_SCIPY ? 0.2 : pb->qmin);
                    rejected = 1;
                    rej++;
                    continue;
                }
                int accept;
                double hn;
                if (pb->controller == CTRL_SCIPY) {
                    accept = err < 1 || forced;
                    if (accept) {
                        double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, expo));
                        if (rejected) factor = fmin(1.0, factor);
                        hn = h * factor;
                    } else
                        hn = h * fmax(0.2, 0.9 * pow(err, expo));
                } else {
                    double q;
                    if (err == 0)
                        q = 1.0 / pb->qmax;
                    else {
                        q11 = pow(err, pb->beta1);
                        q = q11 / pow(qold, pb->beta2);
                        q = fmax(1.0 / pb->qmax, fmin(1.0 / pb->qmin, q / pb->gamma));
                    }
                    accept = err <= 1 || forced;
                    hn = accept ? h / q : h / fmin(1.0 / pb->qmin, q11 / pb->gamma);
                }
                if (!accept) {
                    h_abs = hn;
                    rejected = 1;
                    rej++;
                    continue;
                }
                if (clipped) {
                    if (hn < h_abs) hn = h_abs;
                } else if (pb->controller != CTRL_SCIPY)
                    qold = fmax(err, pb->qoldinit);
                h_abs = hn;
                double t1 = last ? tlim : t + h;
                for (int i = 0; i < nv; ++i) {
                    yprev[i] = u[i];
                    fprev[i] = K[0][i];
                    u[i] = yn[i];
                    K[0][i] = K[S][i];
                }
                hlast = h;
                rejected = 0;
                acc++;
                /* outputs inside (t, t1): Hermite interpolant, OrdinaryDiffEq's hermite_interpolant */
                while (k < nt && ts[k] < t1) {
                    double th = (ts[k] - (t1 - hlast)) / hlast, uo[4];
                    for (int i = 0; i < nv; ++i) {
                        double dy = u[i] - yprev[i];
                        uo[i] = (1 - th) * yprev[i] + th * u[i] +
                                th * (th - 1) * ((1 - 2 * th) * dy + (th - 1) * hlast * fprev[i] + th * hlast * K[0][i]);
                    }
                    if (nonfinite(uo, nv) || orc_out_of_bounds(kind, uo)) {
                        status = ST_DOMAIN;
                        break;
                    }
                    cb(user, k, uo);
                    ++k;
                }
                t = t1;
            }
        } else
        for (k = 0; k < nt && status == ST_OK; ++k) {
            double tstop = ts[k];
            while (t < tstop) {
                if (!have_f) {
                    if (orc_rhs(kind, par, sg, u, K[0])) {
                        status = ST_DOMAIN;
                        break;
                    }
                    h_abs = initial_step(kind, par, sg, u, K[0], tend - t, err_order, tol);
                    have_f = 1;
                }
                if (attempts >= pb->maxiters) {
                    status = ST_MAXITERS;
                    break;
                }
                attempts++;
                int forced = 0;
                if (h_abs < dtmin) {
                    h_abs = dtmin;
                    forced = 1;
                }
                double h = h_abs;
                int clipped = 0;
                if (t + h >= tstop) {
                    clipped = (tstop - t) < h_abs;
                    h = tstop - t;
                }
                int bad;
                if (IS_FIVE(pb->alg))
                    bad = rk_trial(kind, par, sg, 6, FIVE_A(pb->alg), 6, FIVE_B(pb->alg), h, u, K, yn);
                else
                    bad = rk_trial(kind, par, sg, 12, &DOP853_A[0][0], 12, DOP853_B, h, u, K, yn);
                double err;
                {
                    double sc[4];
                    for (int i = 0; i < nv; ++i) sc[i] = tol + fmax(fabs(u[i]), fabs(yn[i])) * tol;
                    if (IS_FIVE(pb->alg)) {
                        double e[4];
                        for (int i = 0; i < nv; ++i) {
                            double s = 0;
                            for (int j = 0; j <= S; ++j) s += FIVE_E(pb->alg)[j] * K[j][i];
                            e[i] = s * h / sc[i];
                        }
                        err = rms(e, nv);
                    } else {
                        double n5 = 0, n3 = 0;
                        for (int i = 0; i < nv; ++i) {
                            double s5 = 0, s3 = 0;
                            for (int j = 0; j <= S; ++j) {
                                s5 += DOP853_E5[j] * K[j][i];
                                s3 += DOP853_E3[j] * K[j][i];
                            }
                            s5 /= sc[i];
                            s3 /= sc[i];
                            n5 += s5 * s5;
                            n3 += s3 * s3;
                        }
                        if (n5 == 0 && n3 == 0)
                            err = 0;
                        else
                            err = fabs(h) * n5 / sqrt((n5 + 0.01 * n3) * nv);
                    }
                }
                bad = bad || nonfinite(yn, nv) || !isfinite(err) || orc_out_of_bounds(kind, yn);
                if (bad) {
                    if (forced) {
                        status = ST_DOMAIN;
                        break;
                    }
                    h_abs = h * (pb->controller == CTRL_SCIPY ? 0.2 : pb->qmin); /* isoutofdomain rejection: shrink by qmin */
                    rejected = 1;
                    rej++;
                    continue;
                }
                if (pb->controller == CTRL_SCIPY) {
                    if (err < 1 || forced) {
                        double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, expo));
                        if (rejected) factor = fmin(1.0, factor);
                        double hn = h * factor;
                        if (clipped && hn < h_abs) hn = h_abs;
                        h_abs = hn;
                        t = (t + h >= tstop) ? tstop : t + h;
                        for (int i = 0; i < nv; ++i) {
                            u[i] = yn[i];
                            K[0][i] = K[S][i];
                        }
                        rejected = 0;
                        acc++;
                    } else {
                        h_abs = h * fmax(0.2, 0.9 * pow(err, expo));
                        rejected = 1;
                        rej++;
                    }
                } else {
                    /* OrdinaryDiffEq: stepsize_controller! / step_accept_controller! / step_reject_controller! */
                    double q;
                    if (err == 0)
                        q = 1.0 / pb->qmax;
                    else {
                        q11 = pow(err, pb->beta1);
                        q = q11 / pow(qold, pb->beta2);
                        q = fmax(1.0 / pb->qmax, fmin(1.0 / pb->qmin, q / pb->gamma));
                    }
                    if (err <= 1 || forced) {
                        double hn = h / q;
                        if (clipped) {
                            /* a step cut short to land on ts[k] (the reference interpolates there instead, src/TWA.jl:180)
                             * does not feed the controller: the un-clipped proposal stands unless the new one is larger */
                            if (hn < h_abs) hn = h_abs;
                        } else
                            qold = fmax(err, pb->qoldinit);
                        h_abs = hn;
                        t = (t + h >= tstop) ? tstop : t + h;
                        for (int i = 0; i < nv; ++i) {
                            u[i] = yn[i];
                            K[0][i] = K[S][i];
                        }
                        acc++;
                    } else {
                        h_abs = h / fmin(1.0 / pb->qmin, q11 / pb->gamma);
                        rej++;
                    }
                }
            }
            if (status != ST_OK) break;
            cb(user, k, u);
        }
    }
    *fail_index = k;
    *nacc = acc;
    *nrej = rej;
    return status;
}

/* dense-output driver used by the parity tests: out[n][nt][nvar] (NaN after a failure) */
typedef struct {
    double *out;
    int nv;
} dense_ctx;
static void dense_cb(void *user, int k, const double *u)
{
    dense_ctx *c = (dense_ctx *)user;
    memcpy(c->out + (size_t)k * c->nv, u, sizeof(double) * c->nv);
}

int orc_integrate(const orc_problem *pb, const double *u0, long n, const double *ts, int nt, double *out,
                  int *status, long *nacc, long *nrej)
{
    int nv = orc_nvar(pb->kind);
    int *nsub = (int *)malloc(sizeof(int) * nt);
    double *hsub = (double *)malloc(sizeof(double) * nt);
    if (orc_step_table(ts, nt, pb->dt > 0 ? pb->dt : 1.0, nsub, hsub)) {
        free(nsub);
        free(hsub);
        return -1;
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < n; ++i) {
        double *o = out + (size_t)i * nt * nv;
        for (long j = 0; j < (long)nt * nv; ++j) o[j] = NAN;
        dense_ctx c = {o, nv};
        int fi;
        long a, r;
        int st = orc_solve_one(pb, u0 + i * nv, ts, nt, nsub, hsub, dense_cb, &c, &fi, &a, &r);
        if (status) status[i] = st;
        if (nacc) nacc[i] = a;
        if (nrej) nrej[i] = r;
    }
    free(nsub);
    free(hsub);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), the counter-based generator the library fuses into its
 * kernels.  Not in the reference (it uses Julia's global MersenneTwister through Distributions,
 * src/TWA.jl:696,740-741); restated here so that samples can be compared bit for bit. */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* two 32-bit words -> a double in (0,1) with 53 random bits, never 0 or 1 */
static double u53(uint32_t a, uint32_t b)
{
    uint64_t m = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

/* the four uniforms of trajectory `idx`, resampling attempt `attempt` */
void orc_uniforms(uint64_t seed, uint64_t idx, uint32_t attempt, double U[4])
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), attempt, 0};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    U[0] = u53(r[0], r[1]);
    U[1] = u53(r[2], r[3]);
    ctr[3] = 1;
    orc_philox4x32_10(ctr, key, r);
    U[2] = u53(r[0], r[1]);
    U[3] = u53(r[2], r[3]);
}

/* Julia's mod(x, 2pi): result carries the sign of the divisor */
static double mod2pi_pos(double x)
{
    double r = fmod(x, TWO_PI);
    if (r < 0) r += TWO_PI;
    return r;
}

/* src/PhaseSpaces.jl:49,63 */
static double theta_of_QP(double Q, double P) { return acos(1 - (Q * Q + P * P) / 2); }
static double phi_of_QP(double Q, double P) { return mod2pi_pos(atan2(-P, Q)); }

typedef struct {
    int kind;         /* SAMPLER_* */
    double center[4]; /* HW: q0,p0 ; SU2: Q0,P0 ; HWxSU2: Q0,q0,P0,p0 */
    double hbar;
    uint64_t seed;
    uint64_t offset; /* global index of the first trajectory */
} orc_sampler;

/* a9: src/TWA.jl:708-745 coherent_Wigner_SU2.sample with Theta ~ Rayleigh(sigma) drawn by inversion
 * from U1 and the azimuth 2*pi*U2; R(theta,phi,theta0,phi0) is src/TWA.jl:712-714 verbatim;
 * Q,P from src/PhaseSpaces.jl:16,34. */
static void sample_su2(double Q0, double P0, double hbar, double U1, double U2, double *Q, double *P)
{
    double sigma = sqrt(hbar / 2);
    double th0 = theta_of_QP(Q0, P0), ph0 = phi_of_QP(Q0, P0);
    double th = sigma * sqrt(-2 * log(U1));
    double ph = U2 * 2 * M_PI;
    double a = cos(th0) * cos(ph) * sin(th) - cos(th) * sin(th0);
    double thr = atan2(sqrt(a * a + sin(th) * sin(th) * sin(ph) * sin(ph)),
                       cos(th) * cos(th0) + cos(ph) * sin(th) * sin(th0));
    double phr = mod2pi_pos(
        atan2(-cos(ph0) * sin(th) * sin(ph) - cos(th0) * cos(ph) * sin(th) * sin(ph0) + cos(th) * sin(th0) * sin(ph0),
              -cos(th0) * cos(ph) * cos(ph0) * sin(th) + cos(th) * cos(ph0) * sin(th0) + sin(th) * sin(ph) * sin(ph0)));
    *Q = sqrt(2 * (1 - cos(thr))) * cos(phr);
    *P = -sqrt(2 * (1 - cos(thr))) * sin(phr);
}

/* a8: src/TWA.jl:689-700 coherent_Wigner_HW.sample, the two normals drawn by Box-Muller */
static void sample_hw(double q0, double p0, double hbar, double U3, double U4, double *q, double *p)
{
    double sigma = sqrt(hbar / 2);
    double r = sigma * sqrt(-2 * log(U3));
    double a = U4 * 2 * M_PI;
    *q = q0 + r * cos(a);
    *p = p0 + r * sin(a);
}

/* a8-a10: one sample, ordering [Q,q,P,p] (src/TWA.jl:758-762) */
void orc_sample_one(const orc_sampler *sp, uint64_t idx, uint32_t attempt, double *u)
{
    double U[4];
    orc_uniforms(sp->seed, idx, attempt, U);
    if (sp->kind == SAMPLER_HW) {
        sample_hw(sp->center[0], sp->center[1], sp->hbar, U[2], U[3], &u[0], &u[1]);
    } else if (sp->kind == SAMPLER_SU2) {
        sample_su2(sp->center[0], sp->center[1], sp->hbar, U[0], U[1], &u[0], &u[1]);
    } else {
        sample_su2(sp->center[0], sp->center[2], sp->hbar, U[0], U[1], &u[0], &u[2]);
        sample_hw(sp->center[1], sp->center[3], sp->hbar, U[2], U[3], &u[1], &u[3]);
    }
}

void orc_sample(const orc_sampler *sp, long n, uint32_t attempt, double *u)
{
    int nv = (sp->kind == SAMPLER_HWXSU2) ? 4 : 2;
    for (long i = 0; i < n; ++i) orc_sample_one(sp, sp->offset + (uint64_t)i, attempt, u + i * nv);
}

/* probability densities: src/TWA.jl:692-694 (MvNormal pdf), :719-736 (W_theta, returning 0.0 where
 * the reference's try/catch swallows an acos DomainError), :757 (product) */
static double pdf_hw(double q0, double p0, double hbar, double q, double p)
{
    double s2 = hbar / 2;
    return exp(-((q - q0) * (q - q0) + (p - p0) * (p - p0)) / (2 * s2)) / (2 * M_PI * s2);
}
static double pdf_su2(double Q0, double P0, double hbar, double Q, double P)
{
    double sigma = sqrt(hbar / 2);
    double a = 1 - (Q * Q + P * P) / 2;
    if (!(a >= -1 && a <= 1)) return 0.0;
    double th = acos(a), ph = phi_of_QP(Q, P);
    double th0 = theta_of_QP(Q0, P0), ph0 = phi_of_QP(Q0, P0);
    double c = cos(th) * cos(th0) + sin(th) * sin(th0) * cos(ph - ph0);
    if (!(c >= -1 && c <= 1)) return 0.0;
    double T = acos(c);
    /* pdf(Normal(0,sigma),T)*sqrt(2pi)*sigma*(1/(2pi sigma^2)) */
    double g = exp(-(T * T) / (2 * sigma * sigma)) / (sqrt(2 * M_PI) * sigma);
    return g * sqrt(2 * M_PI) * sigma * (1 / (2 * M_PI * sigma * sigma));
}
double orc_pdf(const orc_sampler *sp, const double *u)
{
    if (sp->kind == SAMPLER_HW) return pdf_hw(sp->center[0], sp->center[1], sp->hbar, u[0], u[1]);
    if (sp->kind == SAMPLER_SU2) return pdf_su2(sp->center[0], sp->center[1], sp->hbar, u[0], u[1]);
    return pdf_su2(sp->center[0], sp->center[2], sp->hbar, u[0], u[2]) *
           pdf_hw(sp->center[1], sp->center[3], sp->hbar, u[1], u[3]);
}
void orc_pdf_batch(const orc_sampler *sp, const double *u, long n, double *w)
{
    int nv = (sp->kind == SAMPLER_HWXSU2) ? 4 : 2;
    for (long i = 0; i < n; ++i) w[i] = orc_pdf(sp, u + i * nv);
}

/* ------------------------------------------------------------------------------------------ */
/* a17: src/TWA.jl:53-64 scale_to_index.  Returns the 1-based index, or 0 where the reference
 * returns NaN (value outside [start, stop]).  rint = Julia's round (ties to even). */
long orc_scale_to_index(double v, double start, double stop, long len)
{
    if (v < start || stop < v) return 0;
    if (!(v == v)) return 0; /* NaN: the reference would throw InexactError; we drop */
    if (len == 1) return 1;
    double z = ((v - start) / (stop - start)) * (double)(len - 1) + 1;
    if (!(z == z)) return 0; /* degenerate range (start == stop, len > 1): 0/0; the reference would throw InexactError; dropped */
    return (long)rint(z);
}
void orc_scale_to_index_batch(const double *v, long n, double start, double stop, long len, long *idx)
{
    for (long i = 0; i < n; ++i) idx[i] = orc_scale_to_index(v[i], start, stop, len);
}

/* ------------------------------------------------------------------------------------------ */
/* a14/a18/a20: the ensemble driver, src/TWA.jl:100-224, restated with built-in value kinds so
 * that it can be timed without an interpreter in the loop.  General expressions are handled by
 * oracle/oracle.py on orc_integrate's dense output.
 *
 * value kinds:  0 u[a]   1 u[a]^2   2 c*((P^2+Q^2)/2-1) (= c*cos(theta): Weyl.Jz(j)*scale,
 * src/TWA.jl:819,833-835)   3 hamiltonian   4 pdf of the sampling distribution (survival
 * probability, src/TWA.jl:659-661)
 * reducers:     sums[nt][nval] += value (a14, :355-360) and, for every hist spec, counts[nt][len]
 * at scale_to_index(value) (a18 x=:t / y=:t modes, :456-468).
 * Worker split: src/TWA.jl:115-128 with `workers` = OpenMP threads; each worker owns a contiguous
 * range of global trajectory indices; partial results are added (reducer `+`, :163,:369).
 * Error tolerance: src/TWA.jl:186-209 -- a failed trajectory keeps what it contributed before the
 * failure, is resampled (attempt+1), and the replacement contributes from index ti_error on;
 * more than 100 consecutive failures abort. */
typedef struct {
    int kind;
    int a;
    double c;
} orc_value;

typedef struct {
    int value; /* index into the value list */
    double start, stop;
    long len;
} orc_hist;

typedef struct {
    const orc_problem *pb;
    const orc_sampler *sp;
    const orc_value *vals;
    int nval;
    const orc_hist *hists;
    int nhist;
    double *sums;      /* [nt][nval] */
    uint64_t **counts; /* nhist x [nt][len] */
    uint64_t *nvalid;  /* [nt] */
    int ti_error;      /* 0-based */
} ens_ctx;

static double eval_value(const ens_ctx *c, const orc_value *v, const double *u)
{
    int kind = c->pb->kind;
    switch (v->kind) {
    case 0: return u[v->a];
    case 1: return u[v->a] * u[v->a];
    case 2: {
        double Q = u[0], P = (kind == SYS_DICKE) ? u[2] : u[1];
        return v->c * ((P * P + Q * Q) / 2 - 1);
    }
    case 3: return orc_hamiltonian(kind, c->pb->par, u);
    default: return orc_pdf(c->sp, u);
    }
}

static void ens_cb(void *user, int k, const double *u)
{
    ens_ctx *c = (ens_ctx *)user;
    if (k < c->ti_error) return; /* src/TWA.jl:169 */
    double v[16];
    for (int i = 0; i < c->nval; ++i) {
        v[i] = eval_value(c, &c->vals[i], u);
        c->sums[(size_t)k * c->nval + i] += v[i];
    }
    for (int h = 0; h < c->nhist; ++h) {
        const orc_hist *H = &c->hists[h];
        long idx = orc_scale_to_index(v[H->value], H->start, H->stop, H->len);
        if (idx > 0) c->counts[h][(size_t)k * H->len + (idx - 1)]++;
    }
    c->nvalid[k]++;
}

int orc_ensemble(const orc_problem *pb, const orc_sampler *sp, const double *host_u0, long N, const double *ts,
                 int nt, const orc_value *vals, int nval, const orc_hist *hists, int nhist, int workers,
                 int tolerate_errors, double *sums, uint64_t **counts, uint64_t *nvalid, long *steps_acc,
                 long *steps_rej, long *nfailed)
{
    int nv = orc_nvar(pb->kind);
    if (nval > 16) return -2;
    int *nsub = (int *)malloc(sizeof(int) * nt);
    double *hsub = (double *)malloc(sizeof(double) * nt);
    if (orc_step_table(ts, nt, pb->dt > 0 ? pb->dt : 1.0, nsub, hsub)) {
        free(nsub);
        free(hsub);
        return -1;
    }
    if (workers < 1) workers = 1;
    long neff = (N + workers - 1) / workers; /* src/TWA.jl:119 */
    long tot_acc = 0, tot_rej = 0, tot_fail = 0;
    int abort_flag = 0;
    memset(sums, 0, sizeof(double) * (size_t)nt * nval);
    memset(nvalid, 0, sizeof(uint64_t) * nt);
    for (int h = 0; h < nhist; ++h) memset(counts[h], 0, sizeof(uint64_t) * (size_t)nt * hists[h].len);
    ens_ctx *wc = (ens_ctx *)calloc(workers, sizeof(ens_ctx));
#pragma omp parallel for num_threads(workers) schedule(static, 1) reduction(+ : tot_acc, tot_rej, tot_fail)
    for (int w = 0; w < workers; ++w) {
        long lo = (long)w * neff, hi = lo + neff;
        if (hi > N) hi = N;
        if (lo >= hi) continue;
        ens_ctx *c = &wc[w];
        c->pb = pb; c->sp = sp; c->vals = vals; c->nval = nval; c->hists = hists; c->nhist = nhist;
        c->sums = (double *)calloc((size_t)nt * nval, sizeof(double)); /* f_initial, src/TWA.jl:164 */
        c->nvalid = (uint64_t *)calloc(nt, sizeof(uint64_t));
        c->counts = (uint64_t **)malloc(sizeof(uint64_t *) * (nhist > 0 ? nhist : 1));
        for (int h = 0; h < nhist; ++h) c->counts[h] = (uint64_t *)calloc((size_t)nt * hists[h].len, sizeof(uint64_t));
        c->ti_error = 0;
        int nerr = 0;
        uint32_t attempt = 0;
        for (long j = lo; j < hi && !abort_flag;) {
            double u0[4];
            if (host_u0)
                memcpy(u0, host_u0 + j * nv, sizeof(double) * nv);
            else
                orc_sample_one(sp, sp->offset + (uint64_t)j, attempt, u0);
            int fi;
            long a, r;
            int st = orc_solve_one(pb, u0, ts, nt, nsub, hsub, ens_cb, c, &fi, &a, &r);
            tot_acc += a;
            tot_rej += r;
            if (st == ST_OK || st == ST_MAXITERS || host_u0 || !tolerate_errors) {
                if (st != ST_OK) tot_fail++;
                c->ti_error = 0;
                nerr = 0;
                attempt = 0;
                ++j;
            } else {
                tot_fail++;
                if (nerr > 100) { /* src/TWA.jl:202-206 */
                    abort_flag = 1;
                    break;
                }
                nerr++;
                c->ti_error = fi; /* src/TWA.jl:208 */
                attempt++;
            }
        }
    }
    /* the @distributed (+) reduction, src/TWA.jl:163, in worker order so the result is reproducible */
    for (int w = 0; w < workers; ++w) {
        ens_ctx *c = &wc[w];
        if (!c->sums) continue;
        for (size_t i = 0; i < (size_t)nt * nval; ++i) sums[i] += c->sums[i];
        for (int k = 0; k < nt; ++k) nvalid[k] += c->nvalid[k];
        for (int h = 0; h < nhist; ++h)
            for (size_t i = 0; i < (size_t)nt * hists[h].len; ++i) counts[h][i] += c->counts[h][i];
        free(c->sums);
        free(c->nvalid);
        for (int h = 0; h < nhist; ++h) free(c->counts[h]);
        free(c->counts);
    }
    free(wc);
    free(nsub);
    free(hsub);
    *steps_acc = tot_acc;
    *steps_rej = tot_rej;
    *nfailed = tot_fail;
    return abort_flag ? -3 : 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Extended-precision truth: the same equations of motion in x87 long double (64-bit mantissa),
 * fixed-step DOP853 8th-order weights with a small step.  Used only by oracle/make_golden.py to
 * produce tests/golden/ fixtures; itself pinned against 40-digit mpmath Taylor integration there. */
static void rhs_ld(int kind, const double *par, long double sg, const long double *u, long double *du)
{
    if (kind == SYS_DICKE) {
        long double w0 = par[0], w = par[1], g = par[2];
        long double Q = u[0], q = u[1], P = u[2], p = u[3];
        long double s = sqrtl(4 - P * P - Q * Q);
        du[0] = sg * (P * w0 - g * P * q * Q / s);
        du[1] = sg * p * w;
        du[2] = sg * (g * q * (Q * Q) / s - g * q * s - Q * w0);
        du[3] = sg * (-g * Q * s - q * w);
    } else {
        long double Om = par[0], xi = par[1];
        long double Q = u[0], P = u[1];
        du[0] = sg * P * (Om - xi * (Q * Q) / 2);
        du[1] = sg * (Q * (xi * (P * P) / 2 - (2 * xi + Om)) + xi * (Q * Q * Q));
    }
}

void orc_truth_ld(int kind, const double *par, int backwards, const double *u0, long n, const double *ts,
                  int nt, double hmax, double *out)
{
    int nv = orc_nvar(kind);
    long double sg = backwards ? -1.0L : 1.0L;
#pragma omp parallel for schedule(dynamic, 4)
    for (long i = 0; i < n; ++i) {
        long double u[4], K[13][4], yt[4];
        for (int v = 0; v < nv; ++v) u[v] = u0[i * nv + v];
        long double t = 0;
        for (int k = 0; k < nt; ++k) {
            long double d = (long double)ts[k] - t;
            long ns = (long)ceill(d / hmax);
            if (d > 0 && ns < 1) ns = 1;
            long double h = ns ? d / ns : 0;
            for (long s = 0; s < ns; ++s) {
                rhs_ld(kind, par, sg, u, K[0]);
                for (int st = 1; st < 12; ++st) {
                    for (int v = 0; v < nv; ++v) {
                        long double dy = 0;
                        for (int j = 0; j < st; ++j) dy += (long double)DOP853_A[st][j] * K[j][v];
                        yt[v] = u[v] + dy * h;
                    }
                    rhs_ld(kind, par, sg, yt, K[st]);
                }
                for (int v = 0; v < nv; ++v) {
                    long double dy = 0;
                    for (int j = 0; j < 12; ++j) dy += (long double)DOP853_B[j] * K[j][v];
                    u[v] += h * dy;
                }
            }
            t = ts[k];
            for (int v = 0; v < nv; ++v) out[((size_t)i * nt + k) * nv + v] = (double)u[v];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* the variational system and Lyapunov spectra.
 * src/ClassicalSystems.jl:110-128: with get_fundamental_matrix=true the reference integrates the
 * ArrayPartition (x, Phi) with d Phi/dt = J(x) Phi, J = step(system).jac (ParameterizedFunctions'
 * symbolic Jacobian of src/ClassicalDicke.jl:11-17 / src/ClassicalLMG.jl:8-11; written out here).
 * Flat layout y = [x (nv) | Phi column-major (nv*nv)], N = nv + nv*nv <= 20. */
#define VAR_MAXN 20

static void orc_jacobian(int kind, const double *par, double sg, const double *u, double *J /* row-major [nv][nv] */)
{
    if (kind == SYS_DICKE) {
        double w0 = par[0], w = par[1], g = par[2];
        double Q = u[0], q = u[1], P = u[2];
        double s = sqrt(4 - P * P - Q * Q), s3 = s * s * s;
        /* f = g q Q / s */
        double f = g * q * Q / s;
        double fQ = g * q * (1 / s + Q * Q / s3), fq = g * Q / s, fP = g * q * Q * P / s3;
        double sQ = -Q / s, sP = -P / s;
        memset(J, 0, sizeof(double) * 16);
        /* dQ = P w0 - P f */
        J[0 * 4 + 0] = -P * fQ;
        J[0 * 4 + 1] = -P * fq;
        J[0 * 4 + 2] = w0 - f - P * fP;
        /* dq = w p */
        J[1 * 4 + 3] = w;
        /* dP = Q f - g q s - Q w0 */
        J[2 * 4 + 0] = f + Q * fQ - g * q * sQ - w0;
        J[2 * 4 + 1] = Q * fq - g * s;
        J[2 * 4 + 2] = Q * fP - g * q * sP;
        /* dp = -g Q s - q w */
        J[3 * 4 + 0] = -g * s - g * Q * sQ;
        J[3 * 4 + 1] = -w;
        J[3 * 4 + 2] = -g * Q * sP;
        for (int i = 0; i < 16; ++i) J[i] *= sg;
    } else {
        double Om = par[0], xi = par[1];
        double Q = u[0], P = u[1];
        J[0] = -xi * Q * P;
        J[1] = Om - xi * Q * Q / 2;
        J[2] = xi * P * P / 2 - (2 * xi + Om) + 3 * xi * Q * Q;
        J[3] = xi * Q * P;
        for (int i = 0; i < 4; ++i) J[i] *= sg;
    }
}

/* the reference's `sistemaconmatriz` (src/ClassicalSystems.jl:121-125) */
static int orc_var_rhs(int kind, const double *par, double sg, const double *y, double *dy)
{
    int nv = orc_nvar(kind);
    double J[16];
    int bad = orc_rhs(kind, par, sg, y, dy);
    orc_jacobian(kind, par, sg, y, J);
    for (int c = 0; c < nv; ++c)
        for (int r = 0; r < nv; ++r) {
            double s = 0;
            for (int m = 0; m < nv; ++m) s += J[r * nv + m] * y[nv + c * nv + m];
            dy[nv + c * nv + r] = s;
        }
    return bad;
}

static int var_nonfinite(const double *y, int N)
{
    for (int i = 0; i < N; ++i)
        if (!isfinite(y[i])) return 1;
    return 0;
}

static int var_rk_trial(int kind, const double *par, double sg, int N, int S, const double *A, int lda, const double *B,
                        double h, const double *y, double K[][VAR_MAXN], double *ynew)
{
    int bad = 0;
    double yt[VAR_MAXN];
    for (int s = 1; s < S; ++s) {
        for (int i = 0; i < N; ++i) {
            double dy = 0;
            for (int j = 0; j < s; ++j) dy += A[s * lda + j] * K[j][i];
            yt[i] = y[i] + dy * h;
        }
        bad |= orc_var_rhs(kind, par, sg, yt, K[s]);
    }
    for (int i = 0; i < N; ++i) {
        double dy = 0;
        for (int j = 0; j < S; ++j) dy += B[j] * K[j][i];
        ynew[i] = y[i] + h * dy;
    }
    bad |= orc_var_rhs(kind, par, sg, ynew, K[S]);
    return bad;
}

static double var_initial_step(int kind, const double *par, double sg, int N, const double *y0, const double *f0,
                               double interval, int err_order, double tol)
{
    double sc[VAR_MAXN], a[VAR_MAXN], y1[VAR_MAXN], f1[VAR_MAXN];
    if (interval == 0) return 0;
    for (int i = 0; i < N; ++i) sc[i] = tol + fabs(y0[i]) * tol;
    for (int i = 0; i < N; ++i) a[i] = y0[i] / sc[i];
    double d0 = rms(a, N);
    for (int i = 0; i < N; ++i) a[i] = f0[i] / sc[i];
    double d1 = rms(a, N);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > interval) h0 = interval;
    for (int i = 0; i < N; ++i) y1[i] = y0[i] + h0 * f0[i];
    orc_var_rhs(kind, par, sg, y1, f1);
    for (int i = 0; i < N; ++i) a[i] = (f1[i] - f0[i]) / sc[i];
    double d2 = rms(a, N) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15)
        h1 = fmax(1e-6, h0 * 1e-3);
    else
        h1 = pow(0.01 / fmax(d1, d2), 1.0 / (err_order + 1));
    double h = fmin(100 * h0, h1);
    if (!(h == h)) h = 1e-6;
    return fmin(h, interval);
}

/* One (x, Phi) trajectory, same solver contract and controller as orc_solve_one, stepping exactly onto
 * every ts[k]; out[nt][N] (may be NULL: only the final state is kept in y). */
static int var_solve_one(const orc_problem *pb, double *y, const double *ts, int nt, const int *nsub, const double *hsub,
                         double *out, long *nsteps)
{
    int kind = pb->kind, nv = orc_nvar(kind), N = nv + nv * nv;
    double sg = pb->backwards ? -1.0 : 1.0;
    const double *par = pb->par;
    long att = 0;
    int status = ST_OK;
    *nsteps = 0;
    if (var_nonfinite(y, N) || orc_out_of_bounds(kind, y)) return ST_BAD_IC;
    if (pb->alg == ALG_RK4 || pb->alg == ALG_RK8) {
        double K[13][VAR_MAXN], yt[VAR_MAXN], yn[VAR_MAXN];
        for (int k = 0; k < nt; ++k) {
            int bad = 0;
            for (int s = 0; s < nsub[k]; ++s) {
                double h = hsub[k];
                if (pb->alg == ALG_RK4) {
                    bad |= orc_var_rhs(kind, par, sg, y, K[0]);
                    for (int i = 0; i < N; ++i) yt[i] = y[i] + (h / 2) * K[0][i];
                    bad |= orc_var_rhs(kind, par, sg, yt, K[1]);
                    for (int i = 0; i < N; ++i) yt[i] = y[i] + (h / 2) * K[1][i];
                    bad |= orc_var_rhs(kind, par, sg, yt, K[2]);
                    for (int i = 0; i < N; ++i) yt[i] = y[i] + h * K[2][i];
                    bad |= orc_var_rhs(kind, par, sg, yt, K[3]);
                    for (int i = 0; i < N; ++i) y[i] = y[i] + (h / 6) * (K[0][i] + 2 * K[1][i] + 2 * K[2][i] + K[3][i]);
                } else {
                    bad |= orc_var_rhs(kind, par, sg, y, K[0]);
                    bad |= var_rk_trial(kind, par, sg, N, 12, &DOP853_A[0][0], 12, DOP853_B, h, y, K, yn);
                    memcpy(y, yn, sizeof(double) * N);
                }
                att++;
            }
            if (bad || var_nonfinite(y, N) || orc_out_of_bounds(kind, y)) {
                status = ST_DOMAIN;
                break;
            }
            if (out) memcpy(out + (size_t)k * N, y, sizeof(double) * N);
        }
    } else {
        int S = (IS_FIVE(pb->alg)) ? 6 : 12;
        int err_order = (IS_FIVE(pb->alg)) ? 4 : 7;
        double expo = -1.0 / (err_order + 1);
        double tol = pb->tol, dtmin = pb->dtmin;
        double K[13][VAR_MAXN], yn[VAR_MAXN];
        double t = 0, tend = ts[nt - 1], h_abs = 0;
        int have_f = 0, rejected = 0;
        for (int k = 0; k < nt && status == ST_OK; ++k) {
            double tstop = ts[k];
            while (t < tstop) {
                if (!have_f) {
                    if (orc_var_rhs(kind, par, sg, y, K[0])) {
                        status = ST_DOMAIN;
                        break;
                    }
                    h_abs = var_initial_step(kind, par, sg, N, y, K[0], tend - t, err_order, tol);
                    have_f = 1;
                }
                if (att >= pb->maxiters) {
                    status = ST_MAXITERS;
                    break;
                }
                att++;
                int forced = 0;
                if (h_abs < dtmin) {
                    h_abs = dtmin;
                    forced = 1;
                }
                double h = h_abs;
                int clipped = 0;
                if (t + h >= tstop) {
                    clipped = (tstop - t) < h_abs;
                    h = tstop - t;
                }
                int bad;
                if (IS_FIVE(pb->alg))
                    bad = var_rk_trial(kind, par, sg, N, 6, FIVE_A(pb->alg), 6, FIVE_B(pb->alg), h, y, K, yn);
                else
                    bad = var_rk_trial(kind, par, sg, N, 12, &DOP853_A[0][0], 12, DOP853_B, h, y, K, yn);
                double err, sc[VAR_MAXN];
                for (int i = 0; i < N; ++i) sc[i] = tol + fmax(fabs(y[i]), fabs(yn[i])) * tol;
                if (IS_FIVE(pb->alg)) {
                    double e[VAR_MAXN];
                    for (int i = 0; i < N; ++i) {
                        double s = 0;
                        for (int j = 0; j <= S; ++j) s += FIVE_E(pb->alg)[j] * K[j][i];
                        e[i] = s * h / sc[i];
                    }
                    err = rms(e, N);
                } else {
                    double n5 = 0, n3 = 0;
                    for (int i = 0; i < N; ++i) {
                        double s5 = 0, s3 = 0;
                        for (int j = 0; j <= S; ++j) {
                            s5 += DOP853_E5[j] * K[j][i];
                            s3 += DOP853_E3[j] * K[j][i];
                        }
                        s5 /= sc[i];
                        s3 /= sc[i];
                        n5 += s5 * s5;
                        n3 += s3 * s3;
                    }
                    err = (n5 == 0 && n3 == 0) ? 0 : fabs(h) * n5 / sqrt((n5 + 0.01 * n3) * N);
                }
                bad = bad || var_nonfinite(yn, N) || !isfinite(err) || orc_out_of_bounds(kind, yn);
                if (bad) {
                    if (forced) {
                        status = ST_DOMAIN;
                        break;
                    }
                    h_abs = h * 0.2;
                    rejected = 1;
                    continue;
                }
                if (err < 1 || forced) {
                    double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, expo));
                    if (rejected) factor = fmin(1.0, factor);
                    double hn = h * factor;
                    if (clipped && hn < h_abs) hn = h_abs;
                    h_abs = hn;
                    t = (t + h >= tstop) ? tstop : t + h;
                    memcpy(y, yn, sizeof(double) * N);
                    memcpy(K[0], K[S], sizeof(double) * N);
                    rejected = 0;
                } else {
                    h_abs = h * fmax(0.2, 0.9 * pow(err, expo));
                    rejected = 1;
                }
            }
            if (status != ST_OK) break;
            if (out) memcpy(out + (size_t)k * N, y, sizeof(double) * N);
        }
    }
    *nsteps = att;
    return status;
}

/* integrate(...; get_fundamental_matrix=true): out_u[n][nt][nv], out_fm[n][nt][nv*nv] column-major */
int orc_integrate_fm(const orc_problem *pb, const double *u0, const double *fm0, long n, const double *ts, int nt,
                     double *out_u, double *out_fm, int *status, long *nsteps)
{
    int nv = orc_nvar(pb->kind), nn = nv * nv, N = nv + nn;
    int *nsub = (int *)malloc(sizeof(int) * nt);
    double *hsub = (double *)malloc(sizeof(double) * nt);
    if (orc_step_table(ts, nt, pb->dt > 0 ? pb->dt : 1.0, nsub, hsub)) {
        free(nsub);
        free(hsub);
        return -1;
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (long i = 0; i < n; ++i) {
        double y[VAR_MAXN];
        double *tmp = (double *)malloc(sizeof(double) * (size_t)nt * N);
        for (long j = 0; j < (long)nt * N; ++j) tmp[j] = NAN;
        memcpy(y, u0 + i * nv, sizeof(double) * nv);
        for (int c = 0; c < nv; ++c)
            for (int r = 0; r < nv; ++r) y[nv + c * nv + r] = fm0 ? fm0[i * nn + c * nv + r] : (r == c ? 1.0 : 0.0);
        long ns;
        int st = var_solve_one(pb, y, ts, nt, nsub, hsub, tmp, &ns);
        for (int k = 0; k < nt; ++k) {
            memcpy(out_u + ((size_t)i * nt + k) * nv, tmp + (size_t)k * N, sizeof(double) * nv);
            memcpy(out_fm + ((size_t)i * nt + k) * nn, tmp + (size_t)k * N + nv, sizeof(double) * nn);
        }
        free(tmp);
        if (status) status[i] = st;
        if (nsteps) nsteps[i] = ns;
    }
    free(nsub);
    free(hsub);
    return 0;
}

/* lyapunov_spectrum, src/ClassicalSystems.jl:183-224, loop for loop (including the convergence test
 * inside the loop over columns).  status: ST_* of the integration, or 5 = "Maximum iterations". */
int orc_lyapunov(const orc_problem *pb, const double *u0, long n, double t_interval, double ltol, double lreltol,
                 int maxiters, double *lambda, double *x_final, int *iters, int *status, long *nsteps)
{
    int nv = orc_nvar(pb->kind);
    int nsub[1];
    double hsub[1], ts1[1] = {t_interval};
    if (orc_step_table(ts1, 1, pb->dt > 0 ? pb->dt : 1.0, nsub, hsub)) return -1;
#pragma omp parallel for schedule(dynamic, 1)
    for (long ic = 0; ic < n; ++ic) {
        double x[4], u[16], v[16], lam[4] = {0, 0, 0, 0}, lam_old[4], sums[4] = {0, 0, 0, 0};
        memcpy(x, u0 + ic * nv, sizeof(double) * nv);
        for (int c = 0; c < nv; ++c)
            for (int r = 0; r < nv; ++r) u[c * nv + r] = (r == c) ? 1.0 : 0.0;
        int st = ST_OK, done = 0, it = 0;
        long steps = 0;
        for (int k = 1; k <= maxiters && !done && st == ST_OK; ++k) {
            it = k;
            memcpy(lam_old, lam, sizeof(lam));
            double y[VAR_MAXN];
            memcpy(y, x, sizeof(double) * nv);
            memcpy(y + nv, u, sizeof(double) * nv * nv);
            long ns;
            st = var_solve_one(pb, y, ts1, 1, nsub, hsub, NULL, &ns);
            steps += ns;
            if (st != ST_OK) break;
            memcpy(x, y, sizeof(double) * nv);
            for (int i = 0; i < nv && !done; ++i) {
                for (int r = 0; r < nv; ++r) v[i * nv + r] = y[nv + i * nv + r];
                for (int j = 0; j < i; ++j) {
                    double d = 0;
                    for (int r = 0; r < nv; ++r) d += v[i * nv + r] * u[j * nv + r];
                    for (int r = 0; r < nv; ++r) v[i * nv + r] -= d * u[j * nv + r];
                }
                double n2 = 0;
                for (int r = 0; r < nv; ++r) n2 += v[i * nv + r] * v[i * nv + r];
                double nrm = sqrt(n2);
                for (int r = 0; r < nv; ++r) u[i * nv + r] = v[i * nv + r] / nrm;
                sums[i] += log(nrm);
                lam[i] = sums[i] / (k * t_interval);
                double e2 = 0, l2 = 0;
                for (int r = 0; r < nv; ++r) {
                    e2 += (lam_old[r] - lam[r]) * (lam_old[r] - lam[r]);
                    l2 += lam[r] * lam[r];
                }
                if (sqrt(e2) < lreltol * sqrt(l2) + ltol) done = 1;
            }
            /* columns not reached this round keep last round's orthonormal u (they are re-read from y next round only
             * if the loop completed; when `done` the function returns, as in the reference) */
        }
        if (st == ST_OK && !done) st = 5;
        memcpy(lambda + ic * nv, lam, sizeof(double) * nv);
        if (x_final) memcpy(x_final + ic * nv, x, sizeof(double) * nv);
        iters[ic] = it;
        status[ic] = st;
        if (nsteps) nsteps[ic] = steps;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* surfaces of section.  The docs (docs/src/ClassicalDickeExamples.md:82-109) integrate
 * with ContinuousCallback((x,t,_) -> x[4], save; save_positions=(false,false)): every zero crossing of
 * the condition during the solve calls `save`.  DiffEqCallbacks is not in the image; the crossing is
 * located here with Henon's trick (the event function g = coef.u - offset as independent variable, one
 * step of the integrator's tableau of length -g from the end of the crossing step), which the library
 * does too.  ev_t[n][cap], ev_u[n][cap][nv], ev_dir[n][cap], ev_count[n] (counts beyond cap). */
static double ev_g(const double *coef, double offset, const double *u, int nv)
{
    double g = -offset;
    for (int i = 0; i < nv; ++i) g += coef[i] * u[i];
    return g;
}
static void henon_rhs(int kind, const double *par, double sg, const double *coef, const double *y, double *dy)
{
    int nv = orc_nvar(kind);
    double f[4], gd = 0;
    orc_rhs(kind, par, sg, y, f);
    for (int i = 0; i < nv; ++i) gd += coef[i] * f[i];
    for (int i = 0; i < nv; ++i) dy[i] = f[i] / gd;
    dy[nv] = 1.0 / gd;
}
static void henon_step(int kind, const double *par, double sg, const double *coef, int rk4, double hg, double *y)
{
    int N = orc_nvar(kind) + 1;
    double K[13][5], yt[5], yn[5];
    if (rk4) {
        henon_rhs(kind, par, sg, coef, y, K[0]);
        for (int i = 0; i < N; ++i) yt[i] = y[i] + (hg / 2) * K[0][i];
        henon_rhs(kind, par, sg, coef, yt, K[1]);
        for (int i = 0; i < N; ++i) yt[i] = y[i] + (hg / 2) * K[1][i];
        henon_rhs(kind, par, sg, coef, yt, K[2]);
        for (int i = 0; i < N; ++i) yt[i] = y[i] + hg * K[2][i];
        henon_rhs(kind, par, sg, coef, yt, K[3]);
        for (int i = 0; i < N; ++i) y[i] = y[i] + (hg / 6) * (K[0][i] + 2 * K[1][i] + 2 * K[2][i] + K[3][i]);
    } else {
        henon_rhs(kind, par, sg, coef, y, K[0]);
        for (int s = 1; s < 12; ++s) {
            for (int i = 0; i < N; ++i) {
                double dy = 0;
                for (int j = 0; j < s; ++j) dy += DOP853_A[s][j] * K[j][i];
                yt[i] = y[i] + dy * hg;
            }
            henon_rhs(kind, par, sg, coef, yt, K[s]);
        }
        for (int i = 0; i < N; ++i) {
            double dy = 0;
            for (int j = 0; j < 12; ++j) dy += DOP853_B[j] * K[j][i];
            yn[i] = y[i] + hg * dy;
        }
        memcpy(y, yn, sizeof(double) * N);
    }
}

typedef struct {
    const orc_problem *pb;
    const double *coef;
    double offset, g0;
    int direction, cap, count;
    double *ev_t, *ev_u;
    int *ev_dir;
    double us[4], ts0; /* state and time at the start of the current step */
} ev_ctx;

/* locate one crossing inside the bracket [lo, hi] (times measured from the step start c->ts0; g = glo, ghi at its
 * ends): chord estimate, then Newton on the crossing time -- every iteration re-takes the integrator's own step from
 * the step start -- with bisection whenever Newton leaves the bracket; a last Henon step removes what is left of g */
static void ev_refine(ev_ctx *c, double lo, double hi, double glo, double ghi)
{
    int kind = c->pb->kind, nv = orc_nvar(kind);
    double sgn = c->pb->backwards ? -1.0 : 1.0;
    int rk4 = c->pb->alg == ALG_RK4, up = glo < 0;
    double th = lo + (hi - lo) * (glo / (glo - ghi)), z[4], y[5], g2 = ghi;
    if (ghi == 0) th = hi;
    for (int it = 0; it < 8; ++it) {
        memcpy(z, c->us, sizeof(double) * nv);
        if (rk4)
            rk4_step(kind, c->pb->par, sgn, th, z);
        else
            rk8_fixed_step(kind, c->pb->par, sgn, th, z);
        g2 = ev_g(c->coef, c->offset, z, nv);
        if (it == 7 || fabs(g2) <= 1e-13) break;
        if ((g2 < 0) == (glo < 0)) {
            lo = th;
            glo = g2;
        } else
            hi = th;
        double f[4], gd = 0;
        orc_rhs(kind, c->pb->par, sgn, z, f);
        for (int i = 0; i < nv; ++i) gd += c->coef[i] * f[i];
        double tn = th - g2 / gd;
        if (!(tn > lo) || !(tn < hi)) tn = 0.5 * (lo + hi);
        th = tn;
    }
    memcpy(y, z, sizeof(double) * nv);
    y[nv] = c->ts0 + th;
    if (g2 != 0) {
        henon_step(kind, c->pb->par, sgn, c->coef, 1, -g2, y);   /* |g2| <= 1e-13 in all but degenerate cases: a classical RK4 step is exact to rounding */
        int fin = 1;
        for (int i = 0; i <= nv; ++i) fin = fin && isfinite(y[i]);
        if (!fin) {
            memcpy(y, z, sizeof(double) * nv);
            y[nv] = c->ts0 + th;
        }
    }
    if (c->count < c->cap) {
        c->ev_t[c->count] = y[nv];
        memcpy(c->ev_u + (size_t)c->count * nv, y, sizeof(double) * nv);
        c->ev_dir[c->count] = up ? 1 : -1;
    }
    c->count++;
}

/* after an accepted step from (c->ts0, c->us) to (t1, u).  gd0, gd1: dg/dt at the two ends (adaptive pairs, from
 * the FSAL derivatives; have_gd = 0 for fixed steps, which are short enough not to hide a pair of crossings) */
static void ev_check(ev_ctx *c, const double *u, double t1, int have_gd, double gd0, double gd1)
{
    int kind = c->pb->kind, nv = orc_nvar(kind);
    double g0 = c->g0, g1 = ev_g(c->coef, c->offset, u, nv), hs = t1 - c->ts0;
    int up = g0 < 0 && g1 >= 0, down = g0 > 0 && g1 <= 0;
    if ((up && c->direction >= 0) || (down && c->direction <= 0)) {
        ev_refine(c, 0.0, hs, g0, g1);
    } else if (have_gd && !up && !down && g0 * g1 > 0 && gd0 * gd1 < 0) {
        /* tangency: the cubic Hermite interpolant of g has its extremum inside the step, on the other side of zero */
        double d0 = hs * gd0, d1 = hs * gd1, dl = g1 - g0;
        double c2 = 3 * dl - 2 * d0 - d1, c3 = d0 + d1 - 2 * dl;
        double qa = 3 * c3, qb = 2 * c2, disc = qb * qb - 4 * qa * d0, x = -1;
        if (disc >= 0) {
            double sq = sqrt(disc), q = -0.5 * (qb + (qb >= 0 ? sq : -sq));
            double x1 = (qa != 0) ? q / qa : -1, x2 = (q != 0) ? d0 / q : -1;
            x = (x1 > 0 && x1 < 1) ? x1 : x2;
        }
        if (x > 0 && x < 1) {
            double gx = g0 + x * (d0 + x * (c2 + x * c3));
            if (gx * g0 < 0) {
                double thx = x * hs, z[4];
                double sgn = c->pb->backwards ? -1.0 : 1.0;
                memcpy(z, c->us, sizeof(double) * nv);
                if (c->pb->alg == ALG_RK4)
                    rk4_step(kind, c->pb->par, sgn, thx, z);
                else
                    rk8_fixed_step(kind, c->pb->par, sgn, thx, z);
                double gm = ev_g(c->coef, c->offset, z, nv); /* the integrator's own value at the extremum */
                if (gm * g0 < 0) {
                    int first_up = g0 < 0;
                    if ((first_up && c->direction >= 0) || (!first_up && c->direction <= 0)) ev_refine(c, 0.0, thx, g0, gm);
                    if ((!first_up && c->direction >= 0) || (first_up && c->direction <= 0)) ev_refine(c, thx, hs, gm, g1);
                }
            }
        }
    }
    c->g0 = g1;
}

int orc_integrate_events(const orc_problem *pb, const double *u0, long n, double t_end, const double *coef, double offset,
                         int direction, int cap, double *ev_t, double *ev_u, int *ev_dir, int *ev_count, double *u_end,
                         int *status, long *nsteps)
{
    int kind = pb->kind, nv = orc_nvar(kind);
    double sg = pb->backwards ? -1.0 : 1.0;
    const double *par = pb->par;
    int nsub[1];
    double hsub[1], ts1[1] = {t_end};
    if (orc_step_table(ts1, 1, pb->dt > 0 ? pb->dt : 1.0, nsub, hsub)) return -1;
#pragma omp parallel for schedule(dynamic, 1)
    for (long ic = 0; ic < n; ++ic) {
        double u[4];
        memcpy(u, u0 + ic * nv, sizeof(double) * nv);
        ev_ctx c = {pb, coef, offset, 0.0, direction, cap, 0, ev_t + (size_t)ic * cap, ev_u + (size_t)ic * cap * nv,
                    ev_dir + (size_t)ic * cap, {0, 0, 0, 0}, 0.0};
        int st = ST_OK;
        long att = 0;
        if (nonfinite(u, nv) || orc_out_of_bounds(kind, u))
            st = ST_BAD_IC;
        else if (t_end > 0) {
            c.g0 = ev_g(coef, offset, u, nv);
            if (pb->alg == ALG_RK4 || pb->alg == ALG_RK8) {
                for (int s = 0; s < nsub[0]; ++s) {
                    memcpy(c.us, u, sizeof(double) * nv);
                    c.ts0 = (s == 0) ? 0.0 : hsub[0] * (double)s;
                    int bad = (pb->alg == ALG_RK4) ? rk4_step(kind, par, sg, hsub[0], u) : rk8_fixed_step(kind, par, sg, hsub[0], u);
                    att++;
                    if (bad || nonfinite(u, nv) || orc_out_of_bounds(kind, u)) {
                        st = ST_DOMAIN;
                        break;
                    }
                    ev_check(&c, u, (s + 1 == nsub[0]) ? t_end : hsub[0] * (double)(s + 1), 0, 0.0, 0.0);
                }
            } else {
                /* the adaptive loop of orc_solve_one with a per-accepted-step hook */
                int S = (IS_FIVE(pb->alg)) ? 6 : 12, err_order = (IS_FIVE(pb->alg)) ? 4 : 7;
                double expo = -1.0 / (err_order + 1), tol = pb->tol, dtmin = pb->dtmin;
                double K[13][4], yn[4], t = 0, h_abs;
                int rejected = 0;
                if (orc_rhs(kind, par, sg, u, K[0]))
                    st = ST_DOMAIN;
                else {
                    h_abs = initial_step(kind, par, sg, u, K[0], t_end, err_order, tol);
                    while (t < t_end) {
                        if (att >= pb->maxiters) {
                            st = ST_MAXITERS;
                            break;
                        }
                        att++;
                        int forced = 0;
                        if (h_abs < dtmin) {
                            h_abs = dtmin;
                            forced = 1;
                        }
                        double h = h_abs;
                        int clipped = 0;
                        if (t + h >= t_end) {
                            clipped = (t_end - t) < h_abs;
                            h = t_end - t;
                        }
                        int bad = (IS_FIVE(pb->alg)) ? rk_trial(kind, par, sg, 6, FIVE_A(pb->alg), 6, FIVE_B(pb->alg), h, u, K, yn)
                                                       : rk_trial(kind, par, sg, 12, &DOP853_A[0][0], 12, DOP853_B, h, u, K, yn);
                        double err, sc[4];
                        for (int i = 0; i < nv; ++i) sc[i] = tol + fmax(fabs(u[i]), fabs(yn[i])) * tol;
                        if (IS_FIVE(pb->alg)) {
                            double e[4];
                            for (int i = 0; i < nv; ++i) {
                                double s = 0;
                                for (int j = 0; j <= S; ++j) s += FIVE_E(pb->alg)[j] * K[j][i];
                                e[i] = s * h / sc[i];
                            }
                            err = rms(e, nv);
                        } else {
                            double n5 = 0, n3 = 0;
                            for (int i = 0; i < nv; ++i) {
                                double s5 = 0, s3 = 0;
                                for (int j = 0; j <= S; ++j) {
                                    s5 += DOP853_E5[j] * K[j][i];
                                    s3 += DOP853_E3[j] * K[j][i];
                                }
                                s5 /= sc[i];
                                s3 /= sc[i];
                                n5 += s5 * s5;
                                n3 += s3 * s3;
                            }
                            err = (n5 == 0 && n3 == 0) ? 0 : fabs(h) * n5 / sqrt((n5 + 0.01 * n3) * nv);
                        }
                        bad = bad || nonfinite(yn, nv) || !isfinite(err) || orc_out_of_bounds(kind, yn);
                        if (bad) {
                            if (forced) {
                                st = ST_DOMAIN;
                                break;
                            }
                            h_abs = h * 0.2;
                            rejected = 1;
                            continue;
                        }
                        if (err < 1 || forced) {
                            double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, expo));
                            if (rejected) factor = fmin(1.0, factor);
                            double hn = h * factor;
                            if (clipped && hn < h_abs) hn = h_abs;
                            h_abs = hn;
                            memcpy(c.us, u, sizeof(double) * nv);
                            c.ts0 = t;
                            t = (t + h >= t_end) ? t_end : t + h;
                            double gd0 = 0, gd1 = 0; /* dg/dt at both ends, from the FSAL derivatives */
                            for (int i = 0; i < nv; ++i) {
                                gd0 += coef[i] * K[0][i];
                                gd1 += coef[i] * K[S][i];
                            }
                            for (int i = 0; i < nv; ++i) {
                                u[i] = yn[i];
                                K[0][i] = K[S][i];
                            }
                            rejected = 0;
                            ev_check(&c, u, t, 1, gd0, gd1);
                        } else {
                            h_abs = h * fmax(0.2, 0.9 * pow(err, expo));
                            rejected = 1;
                        }
                    }
                }
            }
        }
        ev_count[ic] = c.count;
        status[ic] = st;
        memcpy(u_end + ic * nv, u, sizeof(double) * nv);
        if (nsteps) nsteps[ic] = att;
    }
    return 0;
}
