"""ClassicalSystems -- the abstract system protocol and `integrate`.

Mirrors src/ClassicalSystems.jl:19-45 (protocol) and :78-140 (integrate) of the reference.  The
solve itself runs on the GPU through dtwa_integrate (include/dicketwa.h); nothing here integrates
on the CPU.  The fundamental matrix (get_fundamental_matrix, :110-128) and lyapunov_spectrum /
lyapunov_exponent (:157-224) run on the GPU too (dtwa_integrate_fm, dtwa_lyapunov), and so does the one
DiffEq callback the docs use: ContinuousCallback with save_positions=(false,false) on a linear condition
(surfaces of section, dtwa_integrate_events).  Out of scope: use_big_numbers,
other DiffEq callbacks.
"""
from __future__ import annotations

import numpy as np

from . import _lib


class ClassicalSystem:
    """src/ClassicalSystems.jl:19 -- subclasses provide parameters/varnames/out_of_bounds and a
    dtwa_system kind."""
    _kind = None

    def parameters(self):
        raise NotImplementedError("Please define parameters() for this system")

    def varnames(self):
        raise NotImplementedError("Please define varnames() for this system")

    def out_of_bounds(self, u):
        raise NotImplementedError("Please define out_of_bounds() for this system")

    def _c_system(self):
        s = _lib.System()
        s.kind = self._kind
        for i, v in enumerate(self.parameters()):
            s.params[i] = float(v)
        return s


def parameters(system):
    return system.parameters()


def varnames(system):
    return system.varnames()


def out_of_bounds(system):
    return lambda u, _=None, t=0: system.out_of_bounds(u)


def eye(n):
    """src/ClassicalSystems.jl: the n x n identity (the reference's default `fundamental_matrix_initial`)"""
    return np.eye(int(n))


def step(system):
    """src/ClassicalSystems.jl:30 -- the equations of motion as `f(du, u, p, t)`, filling du in place like the reference's
    ParameterizedFunction (host formula for inspection and for the Newton algebra of UPOs; trajectories are integrated on
    the GPU, never with this).  `p = [parameters..., σ]` with σ = -1 for backward time (:133-134); p=None uses the system's."""
    vf = system.vector_field

    def f(du, u, p=None, t=0.0):
        pars = system.parameters() if p is None else list(p)[:len(system.parameters())]
        sgn = 1.0 if p is None or len(p) <= len(system.parameters()) else float(list(p)[len(system.parameters())])
        du[:] = sgn * np.asarray(vf(u, pars))
        return du
    return f


def nvars(system, *vars):
    """src/ClassicalSystems.jl:37-45 (1-based like the reference)"""
    d = {v: i + 1 for i, v in enumerate(system.varnames())}
    return [d[v] for v in vars]


# ---- integrator algorithms (markers standing for the OrdinaryDiffEq objects the reference passes) ----
class _Alg:
    """name: the reference's spelling; code: the tableau that runs (DTWA_ALG_*); gains: OrdinaryDiffEq's default
    (beta1, beta2, qmin, qmax) of the NAMED algorithm (alg_utils.jl), None = the library's defaults for `code`."""

    def __init__(self, name, code, fixed, note="", gains=None):
        self.name, self.code, self.fixed, self.note, self.gains = name, code, fixed, note, gains

    def __call__(self):
        return self

    def __repr__(self):
        return f"{self.name}()"

    @property
    def runs_as(self):
        return {_lib.ALG_RK4: "RK4", _lib.ALG_RK8: "RK8 (DOP853 weights, fixed step)", _lib.ALG_DP5: "DP5", _lib.ALG_DOP853: "DOP853",
                _lib.ALG_TSIT5: "Tsit5"}[self.code]


DP8 = _Alg("DP8", _lib.ALG_DOP853, False)
DOP853 = DP8
DP5 = _Alg("DP5", _lib.ALG_DP5, False)
RK4 = _Alg("RK4", _lib.ALG_RK4, True)
RK8 = _Alg("RK8", _lib.ALG_RK8, True)
# The reference default (src/ClassicalSystems.jl:84).  Its coefficients are not available offline
# (hard part 8), so the same-order Hairer pair DOP853 (= DP8()) is run instead -- under the step
# controller OrdinaryDiffEq gives the NAMED algorithm (generic order-p gains beta1 = 7/(10p), beta2 = 2/(5p), qmin 1/5, qmax 10).
_GENERIC8 = (7.0 / 80.0, 2.0 / 40.0, 0.2, 10.0)
_GENERIC5 = (7.0 / 50.0, 2.0 / 25.0, 0.2, 10.0)
TsitPap8 = _Alg("TsitPap8", _lib.ALG_DOP853, False, "the Tsitouras-Papakostas tableau is not available offline", _GENERIC8)
Vern8 = _Alg("Vern8", _lib.ALG_DOP853, False, "Verner's tableau is not available offline", _GENERIC8)
# Tsitouras' 5(4) pair itself (tools/gen_tableaux.py: the published coefficients, checked against the order conditions), under the
# generic order-5 gains OrdinaryDiffEq gives it
Tsit5 = _Alg("Tsit5", _lib.ALG_TSIT5, False, "", _GENERIC5)
_BY_NAME = {a.name.lower(): a for a in (DP8, DP5, RK4, RK8, TsitPap8, Vern8, Tsit5)}
_BY_NAME["dop853"] = DP8
_warned_substitution = set()
last_solver = {}   # what the most recent solver_from_kargs resolved to: requested/actual algorithm, controller, gains


def _resolve_alg(alg):
    if alg is None:
        return TsitPap8
    if isinstance(alg, _Alg):
        return alg
    if isinstance(alg, str) and alg.lower().rstrip("()") in _BY_NAME:
        return _BY_NAME[alg.lower().rstrip("()")]
    raise ValueError(f"integrator_alg={alg!r} is not available; use one of {sorted(_BY_NAME)}")


def solver_from_kargs(kargs: dict) -> _lib.Solver:
    """Turn the keyword arguments the reference forwards integrate -> solve
    (src/ClassicalSystems.jl:78-88,138; src/TWA.jl:190) into a dtwa_solver."""
    k = dict(kargs)
    alg = _resolve_alg(k.pop("integrator_alg", k.pop("alg", None)))
    tol = float(k.pop("tol", 1e-12))
    backwards = bool(k.pop("integate_backwards", k.pop("integrate_backwards", False)))
    dt = k.pop("dt", None)
    adaptive = k.pop("adaptive", None)
    dtmin = k.pop("dtmin", None)
    maxiters = k.pop("maxiters", None)
    controller = k.pop("controller", None)
    chart = k.pop("chart", None)
    output = k.pop("output", None)
    gains = {g: k.pop(g, None) for g in ("beta1", "beta2", "qmin", "qmax", "gamma", "qoldinit")}   # solve()'s own keywords
    for ignored in ("save_everystep", "show_progress", "progress", "verbose", "dense"):
        k.pop(ignored, None)
    for unsupported in ("use_big_numbers", "callback"):   # integrate() handles ContinuousCallback itself
        if k.pop(unsupported, None):
            raise NotImplementedError(f"{unsupported} is outside the GPU hot path (SURVEY.md section 8)")
    for unsupported in ("get_fundamental_matrix", "fundamental_matrix_initial"):   # integrate() handles these itself
        if k.pop(unsupported, None) is not None:
            raise NotImplementedError(f"{unsupported} is only available through ClassicalSystems.integrate / lyapunov_spectrum")
    if k:
        raise TypeError(f"unsupported solver keyword(s): {sorted(k)}")
    code = alg.code
    if adaptive is False and not alg.fixed:
        if code == _lib.ALG_DOP853:
            code = _lib.ALG_RK8
        else:
            raise ValueError(f"adaptive=false is only available for RK4 and the 8th-order tableau, not {alg}")
    fixed = code in (_lib.ALG_RK4, _lib.ALG_RK8)
    if fixed and (dt is None or not dt > 0):
        raise ValueError("fixed-step integration needs dt > 0")
    if alg.note and alg.name not in _warned_substitution:   # once per process and algorithm
        _warned_substitution.add(alg.name)
        import warnings
        warnings.warn(f"integrator_alg={alg.name}(): {alg.note}; the same-order pair {alg.runs_as} runs instead, under "
                      f"OrdinaryDiffEq's step controller for {alg.name} (the algorithm actually run is recorded in "
                      "ClassicalSystems.last_solver / TWA.last_info['algorithm'])", RuntimeWarning, stacklevel=3)
    s = _lib.Solver()
    s.alg = code
    s.backwards = int(backwards)
    s.dt = float(dt) if dt else 0.0
    s.tol = tol
    s.dtmin = float(dtmin) if dtmin else 0.0
    s.maxiters = int(maxiters) if maxiters else 0
    if chart in (None, "qp", "QP", "canonical"):
        s.chart = _lib.CHART_QP
    elif chart in ("bloch", "Bloch", "cartesian"):
        s.chart = _lib.CHART_BLOCH     # opt-in: same equations in (jx,jy,jz,q,p); leaves the reference's step sequence
    else:
        raise ValueError(f"chart={chart!r}: use 'qp' (the reference's coordinates, default) or 'bloch'")
    if output in (None, "step", "tstops"):
        s.output = _lib.OUTPUT_STEP
    elif output in ("hermite", "interpolate"):
        s.output = _lib.OUTPUT_HERMITE   # the reference's semantics: outputs interpolated inside un-clipped steps (src/TWA.jl:180)
    else:
        raise ValueError(f"output={output!r}: use 'step' (land on every output time, default) or 'hermite'")
    if controller in (None, "ode", "OrdinaryDiffEq", "pi", "PI"):
        s.controller = _lib.CTRL_ODE
        b1, b2, qmin, qmax = alg.gains if alg.gains is not None else (0.0, 0.0, 0.0, 0.0)
        if gains["beta1"] is not None or gains["beta2"] is not None:
            # OrdinaryDiffEq: beta2 given -> beta1 default follows from it; beta1 alone keeps the default beta2
            d1, d2 = _default_betas(alg, code)
            b2 = float(gains["beta2"]) if gains["beta2"] is not None else d2
            b1 = float(gains["beta1"]) if gains["beta1"] is not None else _beta1_from_beta2(alg, code, b2)
        s.beta1, s.beta2 = b1, b2
        s.qmin = float(gains["qmin"]) if gains["qmin"] is not None else qmin
        s.qmax = float(gains["qmax"]) if gains["qmax"] is not None else qmax
        s.gamma = float(gains["gamma"]) if gains["gamma"] is not None else 0.0
        s.qoldinit = float(gains["qoldinit"]) if gains["qoldinit"] is not None else 0.0
        ctrl_name = "OrdinaryDiffEq PI"
    elif controller in ("scipy", "SciPy"):
        if any(v is not None for v in gains.values()):
            raise ValueError("beta1/beta2/qmin/qmax/gamma/qoldinit belong to the OrdinaryDiffEq controller, not to controller='scipy'")
        s.controller = _lib.CTRL_SCIPY
        ctrl_name = "SciPy"
    else:
        raise ValueError(f"controller={controller!r}: use 'ode' (OrdinaryDiffEq's, the default) or 'scipy'")
    last_solver.clear()
    last_solver.update(requested=alg.name, algorithm={_lib.ALG_RK4: "RK4", _lib.ALG_RK8: "RK8", _lib.ALG_DP5: "DP5", _lib.ALG_DOP853: "DOP853", _lib.ALG_TSIT5: "Tsit5"}[code],
                       substituted=bool(alg.note), controller=None if fixed else ctrl_name, chart="bloch" if s.chart == _lib.CHART_BLOCH else "qp", output="hermite" if (s.output == _lib.OUTPUT_HERMITE and not fixed) else "step",
                       gains=None if (fixed or s.controller != _lib.CTRL_ODE) else effective_gains(s))
    return s


def _default_betas(alg, code):
    """OrdinaryDiffEq alg_utils.jl beta1_default / beta2_default of the algorithm the caller NAMED"""
    if alg.gains is not None:
        return alg.gains[0], alg.gains[1]
    return (0.2 - 0.03, 0.04) if code == _lib.ALG_DP5 else (0.125, 0.0)


def _beta1_from_beta2(alg, code, beta2):
    if alg.gains is not None:     # generic: beta1_default does not depend on beta2
        return alg.gains[0]
    return 0.2 - 3.0 * beta2 / 4.0 if code == _lib.ALG_DP5 else 0.125 - beta2 / 5.0


def effective_gains(s: _lib.Solver):
    """the gains the library will use for this solver (mirror of make_ctrl in csrc/dtwa_api.cu)"""
    tsit5 = s.alg == _lib.ALG_TSIT5
    dp5 = s.alg == _lib.ALG_DP5 or tsit5
    beta2 = max(0.0, s.beta2) if s.beta1 > 0 else (0.08 if tsit5 else 0.04 if dp5 else 0.0)
    beta1 = s.beta1 if s.beta1 > 0 else (0.14 if tsit5 else 0.2 - 3.0 * beta2 / 4.0 if dp5 else 0.125 - beta2 / 5.0)
    return {"beta1": beta1, "beta2": beta2, "qmin": s.qmin if s.qmin > 0 else (0.2 if dp5 else 1.0 / 3.0),
            "qmax": s.qmax if s.qmax > 0 else (10.0 if dp5 else 6.0), "gamma": s.gamma if s.gamma > 0 else 0.9,
            "qoldinit": s.qoldinit if s.qoldinit > 0 else 1e-4}


class ODESolution:
    """What `integrate` returns: times `.t`, states `.u` (len(t) x nvar), `.retcode`; calling it
    with a time re-integrates to that time on the GPU (the reference interpolates)."""

    def __init__(self, system, u0, t0, t, u, retcode, nsteps, kargs):
        self._system, self._u0, self._t0, self._kargs = system, u0, t0, kargs
        self.t, self.u, self.retcode, self.nsteps = t, u, retcode, nsteps
        self.fundamental_matrix = None   # [len(t)][nvar][nvar] when get_fundamental_matrix=True
        self.solver = dict(last_solver)  # requested / actual algorithm, controller and gains of this solve

    def x(self, i=-1):
        """`(x, Φ)` at saved time i -- the reference's `s.u[i].x` of an ArrayPartition (src/ClassicalSystems.jl:118)"""
        if self.fundamental_matrix is None:
            raise ValueError("integrate was called without get_fundamental_matrix=True")
        return self.u[i], self.fundamental_matrix[i]

    def __call__(self, tq):
        tq = np.atleast_1d(np.asarray(tq, dtype=float))
        order = np.argsort(tq)
        s = integrate(self._system, self._u0, float(tq[order][-1]), t0=self._t0, saveat=tq[order], **self._kargs)
        out = np.empty_like(s.u)
        out[order] = s.u
        return out[0] if out.shape[0] == 1 else out

    def __getitem__(self, i):
        return self.u[i]

    def __len__(self):
        return len(self.t)


class _IntegratorView:
    """What `affect!(integrator)` sees in the reference: `.u` and `.t` at the event."""

    def __init__(self, u, t, direction):
        self.u, self.t, self.direction = u, t, direction


class ContinuousCallback:
    """DiffEqBase.ContinuousCallback as the docs use it (docs/src/ClassicalDickeExamples.md:82-109,180-184):
    `condition(x, t, integrator)` must be an affine function of the state (e.g. `lambda x, t, _: x[3]` for p = 0);
    `affect(integrator)` is called, in time order, for every zero crossing found on the GPU and must not modify
    the integrator (save_positions=(false,false) semantics).  `affect_neg` defaults to `affect`; pass None to ignore
    down-crossings.  `abstol`/`rootfind`/`interp_points` are accepted and ignored: crossings are located by a Henon
    step to integrator accuracy, not by a tolerance-driven root search."""
    _same = object()

    def __init__(self, condition, affect, affect_neg=_same, *, save_positions=(False, False), abstol=None, reltol=None,
                 rootfind=True, interp_points=10, max_events=65536):
        if tuple(save_positions) != (False, False):
            raise NotImplementedError("only save_positions=(false,false) is available on the GPU path")
        self.condition, self.affect = condition, affect
        self.affect_neg = affect if affect_neg is ContinuousCallback._same else affect_neg
        self.max_events = int(max_events)

    def linear_form(self, nvar):
        """coef, offset with condition(x) == coef . x - offset (checked at random points)"""
        g0 = float(self.condition(np.zeros(nvar), 0.0, None))
        coef = np.array([float(self.condition(np.eye(nvar)[i], 0.0, None)) - g0 for i in range(nvar)])
        rng = np.random.default_rng(1)
        for _ in range(3):
            x = rng.standard_normal(nvar)
            if abs(float(self.condition(x, 0.0, None)) - (coef @ x + g0)) > 1e-12 * (1 + np.abs(x).max()):
                raise NotImplementedError("ContinuousCallback condition must be an affine function of the state on the GPU path")
        if self.affect is None and self.affect_neg is None:
            raise ValueError("ContinuousCallback without affect")
        direction = 0 if (self.affect is not None and self.affect_neg is not None) else (1 if self.affect is not None else -1)
        return coef, -g0, direction


_RETCODES = {_lib.TRAJ_OK: "Success", _lib.TRAJ_DOMAIN: "DomainError", _lib.TRAJ_MAXITERS: "MaxIters",
             _lib.TRAJ_BAD_IC: "OutOfBounds"}


def integrate(system, u0, t, *, t0=0.0, tol=1e-12, get_fundamental_matrix=False, integrator_alg=None,
              use_big_numbers=False, integate_backwards=False, fundamental_matrix_initial=None, saveat=None,
              callback=None, **kargs):
    """src/ClassicalSystems.jl:78-140.  Integrates u0 from t0 to t; the states are reported at
    `saveat` (array of times in [t0, t]; default [t0, t])."""
    if t < t0:
        raise ValueError("Use integate_backwards=true instead of negative time")  # :89-91
    u0 = np.asarray(u0, dtype=np.float64).ravel()
    names = system.varnames()
    if len(u0) != len(names):
        raise ValueError(f"u₀ should have length {len(names)}, which is the dimension of the phase space of system")  # :97-99
    if system.out_of_bounds(u0):
        raise ValueError(f"The initial condition u₀={u0.tolist()} is out of bounds")  # :100-102
    if use_big_numbers:
        raise NotImplementedError("BigFloat integration is outside the GPU hot path (SURVEY.md section 8)")
    sol = solver_from_kargs(dict(kargs, tol=tol, integrator_alg=integrator_alg, integate_backwards=integate_backwards))
    ts = np.array([t0, t], dtype=float) if saveat is None else np.atleast_1d(np.asarray(saveat, dtype=float))
    if np.any(ts < t0) or np.any(ts > t) or np.any(np.diff(ts) < 0):
        raise ValueError("saveat must be sorted and lie in [t₀, t]")
    full_kargs = dict(kargs, tol=tol, integrator_alg=integrator_alg, integate_backwards=integate_backwards)
    if callback is not None:   # docs/src/ClassicalDickeExamples.md:82-109
        if not isinstance(callback, ContinuousCallback):
            raise NotImplementedError("only ClassicalSystems.ContinuousCallback is available on the GPU path")
        if get_fundamental_matrix or saveat is not None:
            raise NotImplementedError("callback cannot be combined with get_fundamental_matrix / saveat")
        coef, offset, direction = callback.linear_form(len(u0))
        ev_t, ev_u, ev_dir, cnt, u_end, status, nsteps, _ = _lib.default_context().integrate_events(
            system._c_system(), sol, u0[None, :], t - t0, coef, offset, direction, callback.max_events)
        if cnt[0] > callback.max_events:
            raise RuntimeError(f"{int(cnt[0])} events exceed ContinuousCallback(max_events={callback.max_events})")
        for k in range(int(cnt[0])):
            fn = callback.affect if ev_dir[0, k] > 0 else callback.affect_neg
            fn(_IntegratorView(ev_u[0, k].copy(), float(ev_t[0, k]) + t0, int(ev_dir[0, k])))
        return ODESolution(system, u0, t0, ts, np.stack([u0, u_end[0]]), _RETCODES[int(status[0])], int(nsteps[0]), full_kargs)
    if get_fundamental_matrix:   # src/ClassicalSystems.jl:110-128
        n = len(u0)
        fm0 = None
        if fundamental_matrix_initial is not None:
            fm0 = np.asarray(fundamental_matrix_initial, dtype=np.float64)
            if fm0.shape != (n, n):
                raise ValueError(f"fundamental_matrix_initial should be a {n}x{n} matrix")
            fm0 = fm0[None]
        xs, fm, status, nsteps = _lib.default_context().integrate_fm(system._c_system(), sol, u0[None, :], ts - t0, fm0)
        s = ODESolution(system, u0, t0, ts, xs[0], _RETCODES[int(status[0])], int(nsteps[0]), full_kargs)
        s.fundamental_matrix = fm[0]
        return s
    out, status, nsteps = _lib.default_context().integrate(system._c_system(), sol, u0[None, :], ts - t0)
    return ODESolution(system, u0, t0, ts, out[0], _RETCODES[int(status[0])], int(nsteps[0]),
                       dict(kargs, tol=tol, integrator_alg=integrator_alg, integate_backwards=integate_backwards))


def integrate_batch(system, U0, ts, **kargs):
    """Many initial conditions at once (the K3 kernel): returns (out[n][nt][nvar], status[n], nsteps[n]).
    No counterpart in the reference, which loops over `integrate`.  devices=[...] splits the batch over several GPUs."""
    devices = kargs.pop("devices", None)
    sol = solver_from_kargs(kargs)
    if devices:
        U0 = np.ascontiguousarray(U0, dtype=np.float64).reshape(-1, len(system.varnames()))
        return _lib.split_over_devices(devices, len(U0), lambda c, it: c.integrate(system._c_system(), sol, U0[it], ts))
    return _lib.default_context().integrate(system._c_system(), sol, U0, ts)


def integrate_fm_batch(system, U0, ts, fundamental_matrix_initial=None, **kargs):
    """Many (x, Φ) trajectories at once: (x[n][nt][nvar], Φ[n][nt][nvar][nvar], status[n], nsteps[n])."""
    devices = kargs.pop("devices", None)
    sol = solver_from_kargs(kargs)
    if devices:
        U0 = np.ascontiguousarray(U0, dtype=np.float64).reshape(-1, len(system.varnames()))
        nv = U0.shape[1]
        F0 = None if fundamental_matrix_initial is None else np.asarray(fundamental_matrix_initial, dtype=np.float64).reshape(len(U0), nv, nv)
        return _lib.split_over_devices(devices, len(U0), lambda c, it: c.integrate_fm(system._c_system(), sol, U0[it], ts,
                                                                                      None if F0 is None else F0[it]))
    return _lib.default_context().integrate_fm(system._c_system(), sol, U0, ts, fundamental_matrix_initial)


def poincare_section_batch(system, U0, t, *, coef, offset=0.0, direction=0, max_events=1024, packed=False, **kargs):
    """Crossings of the section coef . u = offset for many initial conditions in one launch (the docs' Poincare
    surface, docs/src/ClassicalDickeExamples.md:82-109, draws one `integrate` per initial condition).
    Returns a dict: t[n][cap], u[n][cap][nvar], direction[n][cap], count[n], u_end, status, nsteps, kernel_ms;
    unused slots are NaN.  packed=True (large batches): t[total], u[total][nvar], direction[total] and offsets[n+1]
    instead -- initial condition i owns the entries offsets[i]:offsets[i+1] -- compacted on the GPU, so that neither the
    transfer nor the host arrays carry the [n][max_events] padding."""
    devices = kargs.pop("devices", None)
    sol = solver_from_kargs(kargs)
    csys = system._c_system()
    if packed:
        if devices:
            U0 = np.ascontiguousarray(U0, dtype=np.float64).reshape(-1, len(system.varnames()))
            n, devices = len(U0), list(devices)[:max(1, len(U0))]
            items = [slice(k, n, len(devices)) for k in range(len(devices))]
            parts = _lib.run_on_devices(devices, [lambda c, it=it: c.integrate_events_packed(csys, sol, U0[it], t, coef, offset, direction, max_events)
                                                   for it in items])
            cnt, status, nsteps = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
            u_end = np.empty_like(U0)
            for it, p in zip(items, parts):
                cnt[it], u_end[it], status[it], nsteps[it] = p[4], p[5], p[6], p[7]
            offsets = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.minimum(cnt, max_events), out=offsets[1:])
            total = int(offsets[-1])
            ev_t, ev_u, ev_dir = np.empty(total), np.empty((total, U0.shape[1])), np.empty(total, dtype=np.int32)
            for it, p in zip(items, parts):      # every GPU's rows go to their places in item order
                loc = p[0]
                rows = np.repeat(np.arange(len(loc) - 1), np.diff(loc))
                dest = offsets[:-1][it][rows] + (np.arange(int(loc[-1])) - loc[:-1][rows])
                ev_t[dest], ev_u[dest], ev_dir[dest] = p[1], p[2], p[3]
            ms = max(p[8] for p in parts)
        else:
            offsets, ev_t, ev_u, ev_dir, cnt, u_end, status, nsteps, ms = _lib.default_context().integrate_events_packed(
                csys, sol, U0, t, coef, offset, direction, max_events)
        return dict(t=ev_t, u=ev_u, direction=ev_dir, offsets=offsets, count=cnt, u_end=u_end, status=status, nsteps=nsteps, kernel_ms=ms)
    if devices:
        U0 = np.ascontiguousarray(U0, dtype=np.float64).reshape(-1, len(system.varnames()))
        ev_t, ev_u, ev_dir, cnt, u_end, status, nsteps, ms = _lib.split_over_devices(
            devices, len(U0), lambda c, it: c.integrate_events(csys, sol, U0[it], t, coef, offset, direction, max_events))
    else:
        ev_t, ev_u, ev_dir, cnt, u_end, status, nsteps, ms = _lib.default_context().integrate_events(
            csys, sol, U0, t, coef, offset, direction, max_events)
    return dict(t=ev_t, u=ev_u, direction=ev_dir, count=cnt, u_end=u_end, status=status, nsteps=nsteps, kernel_ms=ms)


last_lyapunov_info = {}


def lyapunov_spectrum_batch(system, X, t=100, *, λtol=1e-4, λreltol=1e-3, maxiters=100, verbose=True, **kargs):
    """lyapunov_spectrum for many initial conditions in one kernel launch (a Lyapunov map,
    docs/src/ClassicalDickeExamples.md:133-236).  Returns (λ[n][nvar], status[n]); status 0 = converged,
    _lib.LYAP_MAXITERS = "Maximum iterations", 1..3 = the integration failed (DTWA_TRAJ_*).  Details of the
    last call (final points, iterations, step counts, kernel time) are left in `last_lyapunov_info`."""
    kargs.setdefault("maxiters_solver", None)
    solver_maxiters = kargs.pop("maxiters_solver")
    for k in ("fundamental_matrix_initial", "save_everystep", "get_fundamental_matrix"):
        kargs.pop(k, None)
    devices = kargs.pop("devices", None)
    sol = solver_from_kargs(dict(kargs, maxiters=solver_maxiters))
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, len(system.varnames()))
    if devices:   # independent orbits: contiguous blocks on several GPUs at once, no exchange step
        lam, xf, iters, status, nsteps, ms = _lib.split_over_devices(
            devices, len(X), lambda c, it: c.lyapunov(system._c_system(), sol, X[it], t, λtol, λreltol, maxiters))
    else:
        lam, xf, iters, status, nsteps, ms = _lib.default_context().lyapunov(system._c_system(), sol, X, t, λtol, λreltol, maxiters)
    last_lyapunov_info.clear()
    last_lyapunov_info.update(x_final=xf, iterations=iters, status=status, nsteps=nsteps, kernel_ms=ms)
    return lam, status


def lyapunov_spectrum(system, x, t=100, *, λtol=1e-4, λreltol=1e-3, maxiters=100, verbose=True, **kargs):
    """src/ClassicalSystems.jl:183-224 (note: the signature's default λtol is 1e-4 there, :185)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    if len(x) != len(system.varnames()):
        raise ValueError(f"u₀ should have length {len(system.varnames())}, which is the dimension of the phase space of system")
    if system.out_of_bounds(x):
        raise ValueError(f"The initial condition u₀={x.tolist()} is out of bounds")
    callback = kargs.pop("callback", None)
    lam, status = lyapunov_spectrum_batch(system, x[None, :], t, λtol=λtol, λreltol=λreltol, maxiters=maxiters, **kargs)
    if status[0] == _lib.LYAP_MAXITERS:
        raise RuntimeError("Maximum iterations")   # :223
    if status[0] != _lib.TRAJ_OK:
        raise RuntimeError(f"integration failed: {_RETCODES[int(status[0])]}")
    if callback is not None:
        # docs/src/ClassicalDickeExamples.md:196-202 pass the section callback through lyapunov_exponent to collect the
        # orbit's Poincare points.  Here the orbit is integrated once more, without its fundamental matrix, over the
        # same total time (rounds x t) with the event kernel: the same orbit, up to the divergence of a chaotic
        # trajectory under a different step sequence.
        rounds = int(last_lyapunov_info["iterations"][0])
        saved = dict(last_lyapunov_info)
        integrate(system, x, rounds * float(t), callback=callback, **kargs)
        last_lyapunov_info.update(saved)
    return lam[0]


def lyapunov_exponent(system, x, t=100, **kargs):
    """src/ClassicalSystems.jl:166-171"""
    return float(np.max(lyapunov_spectrum(system, x, t, **kargs)))
