"""TWA -- the Truncated Wigner Approximation ensemble, on the GPU.

Mirrors the public surface of src/TWA.jl of the reference: PhaseSpaceDistribution (:22-26),
MonteCarloSystem (:37-51), monte_carlo_integrate (:100-224), mcs_chain (:235-274),
mcs_for_averaging / average (:336-370,565-568), mcs_for_distributions / calculate_distribution
(:413-556), mcs_for_variance / variance (:592-633), mcs_for_survival_probability /
survival_probability (:652-680), coherent_Wigner_HW / _SU2 / _HWxSU2 (:689-787) and the Weyl
sub-module (:789-884).

Where the reference runs Julia closures per trajectory on CPU workers, this module only builds POD
descriptions (system, solver, sampler, reducer programs) and calls dtwa_ensemble
(include/dicketwa.h): sampling, integration, observables and the per-time reduction all happen in
one fused CUDA kernel.  What is left on the host is what the reference's f_final closures do:
divide by N, unwrap scalars, orient the histogram matrices.
"""
from __future__ import annotations

import inspect
import itertools
import math
from fractions import Fraction

import numpy as np

from . import ClassicalSystems, PhaseSpaces, distributed, symbolic
from . import _lib
from .symbolic import Expr

last_info = {}  # counters of the most recent monte_carlo_integrate (steps, failures, kernel time ...)


# ------------------------------------------------------------------------------------------------
# ranges: the reference takes Julia ranges `a:s:b` for ts / xs / ys
class JRange:
    """`start:step:stop` with Julia's semantics for decimal literals: the length is
    floor((stop-start)/step)+1 in exact rational arithmetic of the literals' shortest decimal
    representations, elements are correctly rounded (what Julia's twice-precision ranges give)."""

    def __init__(self, start, step, stop):
        a, s, b = (Fraction(repr(float(x))) if not isinstance(x, (int, Fraction)) else Fraction(x) for x in (start, step, stop))
        if s == 0:
            raise ValueError("range step cannot be zero")
        n = int((b - a) / s) + 1 if (b - a) / s >= 0 else 0
        self._a, self._s, self.len = a, s, max(n, 0)
        self._arr = None
        self.start = float(a)
        self.step = float(s)
        self.stop = float(a + (self.len - 1) * s) if self.len else float(a)

    def __len__(self):
        return self.len

    def __array__(self, dtype=None, copy=None):
        # elements a + k s = (pa qs + k ps qa) / (qa qs): Python's int / int is correctly rounded, so this is float(Fraction)
        # without the rational arithmetic (10 ms for 2001 output times -- more than the kernel of a 10^4-trajectory call);
        # computed once per range (a JRange is immutable) and handed out read-only
        if self._arr is None:
            pa, qa, ps, qs = self._a.numerator, self._a.denominator, self._s.numerator, self._s.denominator
            den, n0, dn = qa * qs, pa * qs, ps * qa
            arr = np.array([(n0 + k * dn) / den for k in range(self.len)], dtype=np.float64)
            arr.flags.writeable = False
            self._arr = arr
        a = self._arr if dtype is None else self._arr.astype(dtype, copy=False)
        return a.copy() if copy else a

    def __iter__(self):
        return iter(np.asarray(self))

    def __getitem__(self, i):
        return np.asarray(self)[i]

    def linrange(self):
        """LinRange(start, stop, len) as passed to scale_to_index (src/TWA.jl:435-436)"""
        return (self.start, self.stop, self.len)

    def __repr__(self):
        return f"{self.start}:{self.step}:{self.stop}"


def jrange(start, step_or_stop, stop=None):
    """jrange(a, s, b) == a:s:b ; jrange(a, b) == a:b"""
    if stop is None:
        return JRange(start, 1, step_or_stop)
    return JRange(start, step_or_stop, stop)


def _as_linrange(r, what):
    if isinstance(r, JRange):
        return r.linrange()
    if isinstance(r, tuple) and len(r) == 3 and float(r[2]) == int(r[2]):
        return (float(r[0]), float(r[1]), int(r[2]))
    a = np.asarray(r, dtype=float)
    if a.ndim != 1 or a.size < 1:
        raise ValueError(f"{what} should be a range object")
    if a.size > 2 and not np.allclose(np.diff(a), (a[-1] - a[0]) / (a.size - 1), rtol=1e-9, atol=1e-12):
        raise ValueError(f"{what} should be a uniformly spaced range")
    return (float(a[0]), float(a[-1]), int(a.size))


# ------------------------------------------------------------------------------------------------
class PhaseSpaceDistribution:
    """src/TWA.jl:22-26: (probability_density, sample, ħ).

    `PhaseSpaceDistribution(probability_density, sample, ħ)` -- the reference's own constructor -- builds a USER-DEFINED
    distribution: `sample()` returns one phase-space point (it is called N times on the host and the points go to the
    GPU as host-supplied initial conditions), `probability_density(u)` is needed only by survival_probability and must be
    traceable into an expression (arithmetic and the functions of dickemodel_jl_b200.symbolic), because it is evaluated
    on the GPU along every trajectory.  Trajectories of a user-defined sampler that fail are not re-drawn on the device
    (the library cannot call back into Python): they are reported in TWA.last_info["n_unfinished"].

    The coherent_Wigner_* constructors below build the library's own distributions: samples come from Philox4x32-10
    keyed by `seed` inside the kernel; every draw advances `offset`, so consecutive ensembles use fresh trajectories,
    like consecutive rand() calls in the reference."""

    def __init__(self, kind, center, ħ, seed=0, u0=None):
        self.user_sample = self.user_pdf = None
        if callable(center):   # (probability_density, sample, ħ)
            self.user_pdf, self.user_sample = kind, center
            kind, center = _lib.SAMPLER_HOST_ARRAY, [0, 0, 0, 0]
        self.kind, self.center, self.ħ, self.seed = kind, [float(c) for c in center], float(ħ), int(seed)
        self.offset = 0
        self.u0 = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64)

    def draw_host(self, n, nv):
        """n points of a user-defined sampler as an [n][nvar] array"""
        u = np.array([np.asarray(self.user_sample(), dtype=np.float64).ravel() for _ in range(int(n))], dtype=np.float64).reshape(int(n), -1)
        if n and u.shape[1] != nv:
            raise ValueError(f"the distribution samples {u.shape[1]} variables but the system has {nv}")
        return np.ascontiguousarray(u.reshape(int(n), nv))

    hbar = property(lambda self: self.ħ)

    def _c_sampler(self, offset=None, u0=None):
        s = _lib.Sampler()
        s.kind = self.kind
        for i, c in enumerate(self.center):
            s.center[i] = c
        s.hbar = self.ħ
        s.seed = self.seed
        s.offset = self.offset if offset is None else int(offset)
        if u0 is not None:
            s.u0 = u0.ctypes.data
        return s

    @property
    def nvar(self):
        if self.user_sample is not None:
            return None   # known once a point has been drawn
        if self.u0 is not None:
            return self.u0.shape[1]
        return 4 if self.kind == _lib.SAMPLER_COHERENT_HWXSU2 else 2

    def sample(self):
        if self.user_sample is not None:
            return np.asarray(self.user_sample(), dtype=np.float64)
        if self.u0 is not None:
            u = self.u0[self.offset % len(self.u0)]
            self.offset += 1
            return u.copy()
        u = _lib.default_context().sample(self._c_sampler(), 1)[0]
        self.offset += 1
        return u

    def sample_many(self, n):
        u = _lib.default_context().sample(self._c_sampler(), n)
        self.offset += n
        return u

    def probability_density(self, u):
        if self.user_sample is not None:
            if self.user_pdf is None:
                raise ValueError("this distribution was built without a probability density")
            return float(self.user_pdf(np.asarray(u, dtype=float)))
        if self.u0 is not None:
            raise ValueError("a distribution given by samples has no probability density")
        return float(_lib.default_context().pdf(self._c_sampler(), np.asarray(u, dtype=float))[0])

    def __eq__(self, o):
        return self is o or (isinstance(o, PhaseSpaceDistribution) and self.u0 is None and o.u0 is None
                             and (self.kind, self.center, self.ħ, self.seed) == (o.kind, o.center, o.ħ, o.seed))

    __hash__ = object.__hash__


_seed_counter = itertools.count(20261019)


def _next_seed(seed):
    return int(seed) if seed is not None else next(_seed_counter)


def set_seed(seed):
    """Restart the default seed sequence (all ranks of a multi-process job must agree on it)."""
    global _seed_counter
    _seed_counter = itertools.count(int(seed))


def _hbar(j, ħ):
    return float(ħ) if ħ is not None else 1.0 / float(j)


def coherent_Wigner_HW(u0=None, *, q0=None, p0=None, j=1, ħ=None, hbar=None, seed=None, **kw):
    """src/TWA.jl:689-700,787.  q₀/p₀ may be given as q0/p0 or **{"q₀":..}."""
    q0 = kw.pop("q₀", q0)
    p0 = kw.pop("p₀", p0)
    if u0 is not None:
        q0, p0 = u0[0], u0[1]
    return PhaseSpaceDistribution(_lib.SAMPLER_COHERENT_HW, [q0, p0, 0, 0], _hbar(j, ħ if ħ is not None else hbar), _next_seed(seed))


def coherent_Wigner_SU2(u0=None, *, Q0=None, P0=None, j=1, ħ=None, hbar=None, seed=None, **kw):
    """src/TWA.jl:708-745,779"""
    Q0 = kw.pop("Q₀", Q0)
    P0 = kw.pop("P₀", P0)
    if u0 is not None:
        Q0, P0 = u0[0], u0[1]
    if 4 - Q0 ** 2 - P0 ** 2 < 0:
        raise ValueError("the centre of an SU(2) coherent state must satisfy Q²+P² <= 4")
    return PhaseSpaceDistribution(_lib.SAMPLER_COHERENT_SU2, [Q0, P0, 0, 0], _hbar(j, ħ if ħ is not None else hbar), _next_seed(seed))


def coherent_Wigner_HWxSU2(u0=None, *, Q0=None, q0=None, P0=None, p0=None, j=1, ħ=None, hbar=None, seed=None, **kw):
    """src/TWA.jl:754-771; sample order [Q,q,P,p]"""
    Q0, q0, P0, p0 = kw.pop("Q₀", Q0), kw.pop("q₀", q0), kw.pop("P₀", P0), kw.pop("p₀", p0)
    if u0 is not None:
        Q0, q0, P0, p0 = u0[0], u0[1], u0[2], u0[3]
    if 4 - Q0 ** 2 - P0 ** 2 < 0:
        raise ValueError("the centre of an SU(2) coherent state must satisfy Q²+P² <= 4")
    return PhaseSpaceDistribution(_lib.SAMPLER_COHERENT_HWXSU2, [Q0, q0, P0, p0], _hbar(j, ħ if ħ is not None else hbar), _next_seed(seed))


def samples_from_array(u0, j=1, ħ=None):
    """Host-supplied initial conditions u0[N][nvar] as a distribution (the parity harness: identical
    initial conditions for the oracle and the GPU).  No counterpart in the reference."""
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    if u0.ndim != 2 or u0.shape[1] not in (2, 4):
        raise ValueError("u0 must be [N][nvar]")
    return PhaseSpaceDistribution(_lib.SAMPLER_HOST_ARRAY, [0, 0, 0, 0], _hbar(j, ħ), 0, u0=u0)


# ------------------------------------------------------------------------------------------------
class MonteCarloSystem:
    """src/TWA.jl:37-51.  The reference stores five closures; here `f_loop` is the list of reducer
    programs the kernel evaluates per trajectory and output time, `f_final` the host-side
    finalisation.  f_initial / reducer are implicit (zeroed device accumulators, `+`)."""

    def __init__(self, f_initial, f_loop, reducer, f_final, N, ts, distribution):
        if ts is None:
            ts = [0.0]
        self.f_initial, self.f_loop, self.reducer, self.f_final = f_initial, list(f_loop), reducer, f_final
        self.N = int(N)
        self.ts = ts
        self.distribution = distribution


def _expr_of(system, observable) -> str:
    """src/TWA.jl:277-299 function_from_expression: function of n args / of one vector / symbol /
    expression -> an expression string over the system's varnames."""
    names = system.varnames()
    if isinstance(observable, Expr):
        return observable.s
    if isinstance(observable, str):
        return observable
    if callable(observable):
        syms = symbolic.symbols(names)
        try:
            npar = len([p for p in inspect.signature(observable).parameters.values()
                        if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.default is p.empty])
        except (TypeError, ValueError):
            npar = 1
        try:
            if npar == len(names):
                out = observable(*syms)
            elif npar == 1:
                out = observable(syms)
            else:
                raise TypeError
        except TypeError as e:
            raise ValueError("You should pass a function that takes a n-vector or n arguments, where n is the dimension "
                             "of the phase space (and that is built from arithmetic and the functions of "
                             "dickemodel_jl_b200.symbolic so that it can be traced into an expression).") from e
        return Expr.wrap(out).s
    if isinstance(observable, (int, float)):
        return Expr.wrap(observable).s
    raise TypeError(f"cannot interpret {observable!r} as an observable")


def function_from_expression(system, expression):
    """src/TWA.jl:277-299: observable -> callable of n arguments (or of one n-vector).  The callable evaluates on the
    GPU (dtwa_eval_expr: the observable VM, one IEEE operation per arithmetic node); it also accepts an array
    [m][nvar] and then returns m values."""
    expr = _expr_of(system, expression).lstrip(":")
    csys = system._c_system()
    nv = len(system.varnames())

    def f(*x):
        u = np.asarray(x[0] if len(x) == 1 else x, dtype=np.float64)
        if u.ndim == 1:
            if u.shape[0] != nv:
                raise ValueError("You should pass a n-vector or n arguments, where n is the dimension of the phase space.")
            return float(_lib.default_context().eval_expr(csys, expr, u[None, :])[0])
        return _lib.default_context().eval_expr(csys, expr, u)
    f.expression = expr
    return f


def functions_from_expressions(system, expressions):
    """src/TWA.jl:300-303"""
    return [function_from_expression(system, s) for s in expressions]


def scale_to_index(v, r):
    """src/TWA.jl:53-64 for a LinRange-like r = (start, stop, len) / JRange / array: the 1-based index of the grid point
    nearest to v (ties to even), NaN outside [start, stop].  Evaluated on the GPU with the kernels' own routine."""
    start, stop, length = _as_linrange(r, "r")
    idx = _lib.default_context().scale_to_index(np.atleast_1d(np.asarray(v, dtype=np.float64)), start, stop, length)
    out = np.where(idx > 0, idx.astype(np.float64), np.nan)
    return out if np.ndim(v) else (float("nan") if idx[0] == 0 else int(idx[0]))


def mcs_for_averaging(system, *, observable, ts=(0.0,), N, distribution):
    """src/TWA.jl:336-370"""
    onevar = not isinstance(observable, (list, tuple, np.ndarray))
    obs = [observable] if onevar else list(observable)
    exprs = [_expr_of(system, o) for o in obs]
    spec = {"kind": _lib.RED_AVERAGE, "exprs": exprs}

    def f_final(res):
        res = res / N
        return res[:, 0] if onevar else res
    return MonteCarloSystem(None, [spec], "+", f_final, N, ts, distribution)


def mcs_for_variance(system, *, observable, N, ts=(0.0,), return_average=False, distribution):
    """src/TWA.jl:592-621: averages [f..., f^2...] and combines them."""
    onevar = not isinstance(observable, (list, tuple, np.ndarray))
    obs = [observable] if onevar else list(observable)
    exprs = [_expr_of(system, o) for o in obs]
    tam = len(exprs)
    spec = {"kind": _lib.RED_AVERAGE, "exprs": exprs + [f"({e})^2" for e in exprs]}

    def f_final(res):
        r = res / N
        var = r[:, tam:] - r[:, :tam] ** 2
        mean = r[:, :tam]
        if onevar:
            var, mean = var[:, 0], mean[:, 0]
        return (var, mean) if return_average else var
    return MonteCarloSystem(None, [spec], "+", f_final, N, ts, distribution)


def mcs_for_survival_probability(system, *, N, ts, distribution):
    """src/TWA.jl:652-668"""
    spec = {"kind": _lib.RED_SURVIVAL}
    if distribution.user_sample is not None:   # user-defined distribution: its density, traced into an expression, is the observable
        if distribution.user_pdf is None:
            raise ValueError("survival_probability needs a distribution with a probability density")
        spec = {"kind": _lib.RED_AVERAGE, "exprs": [_expr_of(system, distribution.user_pdf)]}
    L = len(system.varnames()) / 2

    def f_final(res):
        return res[:, 0] * (2 * math.pi * distribution.ħ) ** L / N
    return MonteCarloSystem(None, [spec], "+", f_final, N, ts, distribution)


def _is_t(a):
    return isinstance(a, str) and a in ("t", ":t")


def mcs_for_distributions(system, *, x, xs=None, y=None, ts=None, animate=False, ys=None, N, distribution):
    """src/TWA.jl:413-543.  Returns matrices in the reference's orientation: [iy, ix]
    (x=:t -> len(ys) x len(ts); y=:t -> len(ts) x len(xs))."""
    if ts is None and y is None:  # :417-420
        y = "t"
        ts = jrange(0, 0)
    if ts is None:
        raise ValueError("You have to pass :ts")  # :421-423
    xt, yt = _is_t(x), _is_t(y)
    if y is None:
        raise ValueError("You have to pass y (an observable or :t)")
    spec = {"kind": _lib.RED_HIST, "x_axis": _lib.AXIS_TIME if xt else _lib.AXIS_EXPR,
            "y_axis": _lib.AXIS_TIME if yt else _lib.AXIS_EXPR, "animate": bool(animate) and not (xt or yt)}
    if not xt:
        if xs is None:
            raise ValueError("xs should be given if x is not :t")
        spec["x_expr"] = _expr_of(system, x)
        spec["xs"] = _as_linrange(xs, "xs")
    if not yt:
        if ys is None:
            raise ValueError("ys should be given if y is not :t")
        spec["y_expr"] = _expr_of(system, y)
        spec["ys"] = _as_linrange(ys, "ys")

    def f_final(counts):
        m = counts.astype(np.float64) / N  # :518
        if xt:
            return m.T  # library [nt][ny] -> reference matrices[iy, i_t]
        if spec["animate"]:
            return [mk for mk in m]  # one [iy, ix] matrix per time (:496,:525)
        return m
    return MonteCarloSystem(None, [spec], "+", f_final, N, ts, distribution)


def mcs_chain(*Fs):
    """src/TWA.jl:235-274"""
    if len(Fs) == 1 and not isinstance(Fs[0], MonteCarloSystem):
        Fs = tuple(Fs[0])
    Fs = list(Fs)
    N, ts, dist = Fs[0].N, Fs[0].ts, Fs[0].distribution
    if any(F.N != N for F in Fs):
        raise ValueError("All systems must have the same N")
    if any(len(np.asarray(F.ts)) != len(np.asarray(ts)) or np.any(np.asarray(F.ts) != np.asarray(ts)) for F in Fs):
        raise ValueError("All systems must have the same ts")
    if any(F.distribution != dist for F in Fs):
        raise ValueError("All systems must have the same phase space distribution")
    sizes = [len(F.f_loop) for F in Fs]

    def f_final(res):
        out, i = [], 0
        for F, n in zip(Fs, sizes):
            chunk = res[i:i + n]
            out.append(F.f_final(chunk if n > 1 else chunk[0]))
            i += n
        return out
    mc = MonteCarloSystem(None, [s for F in Fs for s in F.f_loop], "+", f_final, N, ts, dist)
    mc._chained = True
    return mc


def monte_carlo_integrate(system, mc_system, *, tolerate_errors=True, show_progress=True, maxNBatch=math.inf, **kargs):
    """src/TWA.jl:100-224, on the GPU.  kargs go to the solver (tol, integrator_alg,
    integate_backwards, dt, adaptive, maxiters, dtmin), as they go to ClassicalSystems.integrate in
    the reference (:190).  maxNBatch bounds the trajectories per launch (:119)."""
    global last_info
    for k in ("ts", "N", "distribution"):  # the docs pass these again (docs/src/TWAExamples.md:126-131); they are ignored
        kargs.pop(k, None)
    ctx = _lib.default_context()
    dist = mc_system.distribution
    N = mc_system.N
    ts = np.asarray(mc_system.ts, dtype=np.float64).ravel()
    if ts.size == 0:
        raise ValueError("ts is empty")
    if N < 1:
        raise ValueError("N must be >= 1")
    nv = len(system.varnames())
    if dist.nvar is not None and dist.nvar != nv:
        raise ValueError(f"the distribution samples {dist.nvar} variables but the system has {nv}")
    sol = ClassicalSystems.solver_from_kargs(kargs)
    csys = system._c_system()
    dcfg = _lib.distributed()
    rank, world = dcfg["rank"], dcfg["world"]
    lo, hi = distributed.shard_bounds(N, rank, world)
    # the launch count must not depend on the rank (every launch ends in a collective when world > 1): a rank whose
    # block is shorter, or empty, runs empty chunks
    chunk, nchunks = distributed.chunk_plan(N, world, maxNBatch)
    base = dist.offset
    if dist.u0 is not None and base + N > len(dist.u0):
        raise ValueError("not enough host-supplied initial conditions for N")
    total, info_tot = None, {}
    for c in range(nchunks):
        pos = min(hi, lo + c * chunk)
        n = max(0, min(chunk, hi - pos))
        if dist.user_sample is not None:      # user-defined sampler: n calls on the host, points handed over as an array
            u0 = dist.draw_host(n, nv) if n > 0 else None
        else:
            u0 = dist.u0[base + pos: base + pos + n] if (dist.u0 is not None and n > 0) else None
        smp = dist._c_sampler(offset=base + pos, u0=u0)
        arrays, info = ctx.ensemble(csys, sol, smp, n, ts, mc_system.f_loop, tolerate_errors=tolerate_errors,
                                    allreduce=dcfg["allreduce"] if world > 1 else None)
        if total is None:
            total, info_tot = arrays, info
        else:
            for a, b in zip(total, arrays):
                a += b
            for k, v in info.items():
                info_tot[k] = max(info_tot[k], v) if k in ("resample_rounds", "jit") else info_tot[k] + v
    dist.offset = base + N
    info_tot.update(algorithm=ClassicalSystems.last_solver.get("algorithm"), requested_algorithm=ClassicalSystems.last_solver.get("requested"),
                    controller=ClassicalSystems.last_solver.get("controller"), controller_gains=ClassicalSystems.last_solver.get("gains"))
    if tolerate_errors and info_tot.get("n_unfinished", 0) > 0:
        import warnings
        warnings.warn(f"{info_tot['n_unfinished']} of {N} trajectories did not finish (maxiters reached, or more failures than "
                      "the resampling list holds) and contribute nothing after they stopped; results are still divided by N "
                      "(TWA.last_info['n_valid'] has the per-time counts)", RuntimeWarning, stacklevel=2)
    last_info = info_tot
    res = total if (len(total) > 1 or getattr(mc_system, "_chained", False)) else total[0]
    return mc_system.f_final(res)


def average(system, *, observable, N, ts=(0.0,), distribution, **kargs):
    """src/TWA.jl:565-568"""
    return monte_carlo_integrate(system, mcs_for_averaging(system, observable=observable, ts=ts, N=N, distribution=distribution), **kargs)


def variance(system, *, observable, N, ts=(0.0,), distribution, return_average=False, **kargs):
    """src/TWA.jl:630-633"""
    return monte_carlo_integrate(system, mcs_for_variance(system, observable=observable, ts=ts, N=N, distribution=distribution,
                                                          return_average=return_average), **kargs)


def survival_probability(system, *, distribution, N, ts, **kargs):
    """src/TWA.jl:677-680: trajectories run backwards in time"""
    kargs.setdefault("integate_backwards", True)
    return monte_carlo_integrate(system, mcs_for_survival_probability(system, ts=ts, N=N, distribution=distribution), **kargs)


def calculate_distribution(system, *, x, N, animate=False, xs=None, y=None, ts=None, ys=None, distribution, **kargs):
    """src/TWA.jl:552-556"""
    return monte_carlo_integrate(system, mcs_for_distributions(system, x=x, y=y, ts=ts, xs=xs, ys=ys, N=N,
                                                               distribution=distribution, animate=animate), **kargs)


# ------------------------------------------------------------------------------------------------
class Weyl:
    """src/TWA.jl:789-884: Weyl symbols as expressions in Q,q,P,p (operation order of the source)."""
    @staticmethod
    def _angles():
        Q, P = Expr("Q"), Expr("P")
        cosθ = (P ** 2 + Q ** 2) / 2 - 1
        sinθ = symbolic.sqrt(1 - cosθ ** 2)
        cosφ = Q / symbolic.sqrt(P ** 2 + Q ** 2)
        sinφ = -P / symbolic.sqrt(P ** 2 + Q ** 2)
        return cosθ, sinθ, cosφ, sinφ

    @staticmethod
    def n(j):
        q, p = Expr("q"), Expr("p")
        return (j * (p ** 2 + q ** 2) - 1) / 2

    @staticmethod
    def n2(j):
        return Weyl.n(j) ** 2 - 1 / 4

    @staticmethod
    def Jz(j):
        c, s, cp, sp = Weyl._angles()
        return math.sqrt(j * (j + 1)) * c

    @staticmethod
    def Jx(j):
        c, s, cp, sp = Weyl._angles()
        return math.sqrt(j * (j + 1)) * s * cp

    @staticmethod
    def Jy(j):
        c, s, cp, sp = Weyl._angles()
        return math.sqrt(j * (j + 1)) * s * sp

    @staticmethod
    def _sq(j, inner):
        bj = math.sqrt(j * (j + 1) * (2 * j - 1) * (2 * j + 3))
        return (j * (j + 1)) / 3 + (bj / 2) * (inner ** 2 - 1 / 3)

    @staticmethod
    def Jz2(j):
        return Weyl._sq(j, Weyl._angles()[0])

    @staticmethod
    def Jx2(j):
        c, s, cp, sp = Weyl._angles()
        return Weyl._sq(j, s * cp)

    @staticmethod
    def Jy2(j):
        c, s, cp, sp = Weyl._angles()
        return Weyl._sq(j, s * sp)


for _a, _b in (("n²", "n2"), ("Jz²", "Jz2"), ("Jx²", "Jx2"), ("Jy²", "Jy2")):
    setattr(Weyl, _a, getattr(Weyl, _b))  # reachable as getattr(Weyl, "Jz²"); ² is not a Python identifier character
