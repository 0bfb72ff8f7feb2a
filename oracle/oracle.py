"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/oracle.c plus numpy restatements of the reference's reducers
(/root/reference/src/TWA.jl).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  PARITY UNPINNED -- see the header of oracle.c.

Nothing here reads /root/reference at run time and nothing here touches the GPU.
"""
from __future__ import annotations

import ast
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")

SYS_DICKE, SYS_LMG = 0, 1
ALG_RK4, ALG_RK8, ALG_DP5, ALG_DOP853, ALG_TSIT5 = 0, 1, 2, 3, 4
ST_OK, ST_DOMAIN, ST_MAXITERS, ST_BAD_IC = 0, 1, 2, 3
SAMPLER_HOST, SAMPLER_HW, SAMPLER_SU2, SAMPLER_HWXSU2 = 0, 1, 2, 3


def build(force: bool = False) -> None:
    """Compile oracle.c (both flavours) with the Makefile next to it."""
    need = force or not all(
        os.path.exists(os.path.join(BUILD, n)) for n in ("liboracle.so", "liboracle_fast.so"))
    if not need:
        src = max(os.path.getmtime(os.path.join(HERE, f)) for f in ("oracle.c", "tableaux_gen.h"))
        need = src > min(os.path.getmtime(os.path.join(BUILD, n)) for n in ("liboracle.so", "liboracle_fast.so"))
    if need:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", HERE, "-s"], check=True, env=env)


class Problem(C.Structure):
    _fields_ = [("kind", C.c_int), ("par", C.c_double * 4), ("backwards", C.c_int), ("alg", C.c_int),
                ("dt", C.c_double), ("tol", C.c_double), ("dtmin", C.c_double), ("maxiters", C.c_long),
                ("controller", C.c_int), ("beta1", C.c_double), ("beta2", C.c_double), ("qmin", C.c_double),
                ("qmax", C.c_double), ("gamma", C.c_double), ("qoldinit", C.c_double), ("interp", C.c_int)]


class Sampler(C.Structure):
    _fields_ = [("kind", C.c_int), ("center", C.c_double * 4), ("hbar", C.c_double),
                ("seed", C.c_uint64), ("offset", C.c_uint64)]


class Value(C.Structure):
    _fields_ = [("kind", C.c_int), ("a", C.c_int), ("c", C.c_double)]


class Hist(C.Structure):
    _fields_ = [("value", C.c_int), ("start", C.c_double), ("stop", C.c_double), ("len", C.c_long)]


_libs = {}


def lib(fast: bool = False):
    key = "fast" if fast else "strict"
    if key not in _libs:
        build()
        L = C.CDLL(os.path.join(BUILD, "liboracle_fast.so" if fast else "liboracle.so"))
        dp = C.POINTER(C.c_double)
        L.orc_rhs.argtypes = [C.c_int, dp, C.c_double, dp, dp]
        L.orc_rhs.restype = C.c_int
        L.orc_hamiltonian_batch.argtypes = [C.c_int, dp, dp, C.c_long, dp]
        L.orc_step_table.argtypes = [dp, C.c_int, C.c_double, C.POINTER(C.c_int), dp]
        L.orc_step_table.restype = C.c_int
        L.orc_integrate.argtypes = [C.POINTER(Problem), dp, C.c_long, dp, C.c_int, dp,
                                    C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_integrate.restype = C.c_int
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, dp]
        L.orc_sample.argtypes = [C.POINTER(Sampler), C.c_long, C.c_uint32, dp]
        L.orc_pdf_batch.argtypes = [C.POINTER(Sampler), dp, C.c_long, dp]
        L.orc_scale_to_index_batch.argtypes = [dp, C.c_long, C.c_double, C.c_double, C.c_long,
                                               C.POINTER(C.c_long)]
        L.orc_ensemble.argtypes = [C.POINTER(Problem), C.POINTER(Sampler), dp, C.c_long, dp, C.c_int,
                                   C.POINTER(Value), C.c_int, C.POINTER(Hist), C.c_int, C.c_int, C.c_int,
                                   dp, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.orc_ensemble.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        ip, lp = C.POINTER(C.c_int), C.POINTER(C.c_long)
        L.orc_integrate_fm.argtypes = [C.POINTER(Problem), dp, dp, C.c_long, dp, C.c_int, dp, dp, ip, lp]
        L.orc_integrate_fm.restype = C.c_int
        L.orc_lyapunov.argtypes = [C.POINTER(Problem), dp, C.c_long, C.c_double, C.c_double, C.c_double, C.c_int,
                                   dp, dp, ip, ip, lp]
        L.orc_lyapunov.restype = C.c_int
        L.orc_integrate_events.argtypes = [C.POINTER(Problem), dp, C.c_long, C.c_double, dp, C.c_double, C.c_int, C.c_int,
                                           dp, dp, ip, ip, dp, ip, lp]
        L.orc_integrate_events.restype = C.c_int
        _libs[key] = L
    return _libs[key]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def nvar(kind):
    return 4 if kind == SYS_DICKE else 2


def varnames(kind):
    """src/ClassicalDicke.jl:21, src/ClassicalLMG.jl:15"""
    return ["Q", "q", "P", "p"] if kind == SYS_DICKE else ["Q", "P"]


CTRL_ODE, CTRL_SCIPY = 0, 1
# OrdinaryDiffEq v5 alg_utils.jl defaults (beta1, beta2, qmin, qmax); gamma = 9/10, qoldinit = 1e-4 for every explicit pair:
#   DP5:  beta2 = 4/100, beta1 = 1/5 - 3 beta2/4;  DP8: beta2 = 0, beta1 = 1/8, qmin = 1/3, qmax = 6;
#   anything else of order p (the reference's default TsitPap8, p = 8): beta2 = 2/(5p), beta1 = 7/(10p), qmin = 1/5, qmax = 10
ODE_GAINS = {"dp5": (0.2 - 0.03, 0.04, 0.2, 10.0), "dp8": (0.125, 0.0, 1.0 / 3.0, 6.0), "tsitpap8": (7.0 / 80.0, 2.0 / 40.0, 0.2, 10.0),
             "tsit5": (7.0 / 50.0, 2.0 / 25.0, 0.2, 10.0)}
_controller_default = [None]   # tests may pin a module-wide default with set_default_controller


def set_default_controller(c):
    _controller_default[0] = c


def _controller_fields(p, alg, controller):
    """controller: None/"ode" (OrdinaryDiffEq gains of the algorithm), "scipy", a key of ODE_GAINS, or (beta1, beta2, qmin, qmax)"""
    if controller is None:
        controller = _controller_default[0]
    p.gamma, p.qoldinit = 0.9, 1e-4
    if controller == "scipy":
        p.controller = CTRL_SCIPY
        p.beta1, p.beta2, p.qmin, p.qmax = 0.0, 0.0, 0.2, 10.0
        return
    p.controller = CTRL_ODE
    if controller is None or controller == "ode":
        controller = "dp5" if alg == ALG_DP5 else "tsit5" if alg == ALG_TSIT5 else "dp8"
    if isinstance(controller, str):
        controller = ODE_GAINS[controller.lower()]
    p.beta1, p.beta2, p.qmin, p.qmax = (float(x) for x in controller[:4])
    if len(controller) > 4:
        p.gamma = float(controller[4])


def make_problem(kind, par, alg=ALG_DOP853, dt=0.0, tol=1e-12, dtmin=None, maxiters=100000, backwards=False, controller=None, output=None):
    p = Problem()
    _controller_fields(p, alg, controller)
    p.interp = 1 if output in ("hermite", "interpolate") else 0
    p.kind = kind
    for i, v in enumerate(par):
        p.par[i] = float(v)
    p.backwards = int(bool(backwards))
    p.alg = alg
    p.dt = float(dt)
    p.tol = float(tol)
    p.dtmin = float(tol if dtmin is None else dtmin)  # src/ClassicalSystems.jl:138 dtmin=tol
    p.maxiters = int(maxiters)
    return p


def make_sampler(kind, center, hbar, seed=0, offset=0):
    s = Sampler()
    s.kind = kind
    for i, v in enumerate(center):
        s.center[i] = float(v)
    s.hbar = float(hbar)
    s.seed = int(seed)
    s.offset = int(offset)
    return s


def rhs(kind, par, u, sigma=1.0):
    par = np.ascontiguousarray(par, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    du = np.empty_like(u)
    bad = lib().orc_rhs(kind, _dp(par), float(sigma), _dp(u), _dp(du))
    return du, bool(bad)


def hamiltonian(kind, par, u):
    par = np.ascontiguousarray(par, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nvar(kind))
    h = np.empty(u.shape[0])
    lib().orc_hamiltonian_batch(kind, _dp(par), _dp(u), u.shape[0], _dp(h))
    return h


def step_table(ts, dt):
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    nsub = np.empty(len(ts), dtype=np.int32)
    hsub = np.empty(len(ts))
    rc = lib().orc_step_table(_dp(ts), len(ts), float(dt), nsub.ctypes.data_as(C.POINTER(C.c_int)), _dp(hsub))
    if rc:
        raise ValueError("ts must be sorted and non-negative")
    return nsub, hsub


def integrate(kind, par, u0, ts, alg=ALG_DOP853, dt=0.0, tol=1e-12, dtmin=None, maxiters=100000,
              backwards=False, fast=False, controller=None, output=None):
    """Dense restatement of ClassicalSystems.integrate + the per-time callback: returns
    (out[n][nt][nvar], status[n], accepted[n], rejected[n])."""
    nv = nvar(kind)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    n, nt = u0.shape[0], len(ts)
    out = np.empty((n, nt, nv))
    status = np.empty(n, dtype=np.int32)
    nacc = np.empty(n, dtype=np.int64)
    nrej = np.empty(n, dtype=np.int64)
    pb = make_problem(kind, par, alg, dt, tol, dtmin, maxiters, backwards, controller, output)
    rc = lib(fast).orc_integrate(C.byref(pb), _dp(u0), n, _dp(ts), nt, _dp(out),
                                 status.ctypes.data_as(C.POINTER(C.c_int)),
                                 nacc.ctypes.data_as(C.POINTER(C.c_long)),
                                 nrej.ctypes.data_as(C.POINTER(C.c_long)))
    if rc:
        raise ValueError("ts must be sorted and non-negative")
    return out, status, nacc, nrej


def integrate_fm(kind, par, u0, ts, fm0=None, alg=ALG_DOP853, dt=0.0, tol=1e-12, dtmin=None, maxiters=100000,
                 backwards=False, fast=False):
    """integrate(...; get_fundamental_matrix=true) (src/ClassicalSystems.jl:110-128): returns
    (x[n][nt][nv], Phi[n][nt][nv][nv] with Phi[..., r, c] = d x_r(t) / d x_c(0), status[n], attempts[n])."""
    nv = nvar(kind)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    n, nt = u0.shape[0], len(ts)
    f0 = None
    if fm0 is not None:   # [n][r][c] -> column-major flat
        f0 = np.ascontiguousarray(np.swapaxes(np.asarray(fm0, dtype=np.float64).reshape(n, nv, nv), 1, 2))
    out_u, out_fm = np.empty((n, nt, nv)), np.empty((n, nt, nv * nv))
    status, nsteps = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
    pb = make_problem(kind, par, alg, dt, tol, dtmin, maxiters, backwards)
    rc = lib(fast).orc_integrate_fm(C.byref(pb), _dp(u0), _dp(f0) if f0 is not None else None, n, _dp(ts), nt, _dp(out_u),
                                    _dp(out_fm), status.ctypes.data_as(C.POINTER(C.c_int)),
                                    nsteps.ctypes.data_as(C.POINTER(C.c_long)))
    if rc:
        raise ValueError("ts must be sorted and non-negative")
    return out_u, np.swapaxes(out_fm.reshape(n, nt, nv, nv), 2, 3), status, nsteps


def lyapunov(kind, par, u0, t=100.0, ltol=1e-4, lreltol=1e-3, maxiters=100, alg=ALG_DOP853, dt=0.0, tol=1e-12,
             dtmin=None, solver_maxiters=10 ** 9, fast=False):
    """lyapunov_spectrum (src/ClassicalSystems.jl:183-224) for a batch: (lambda[n][nv], x_final, iters, status, attempts)."""
    nv = nvar(kind)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
    n = u0.shape[0]
    lam, xf = np.empty((n, nv)), np.empty((n, nv))
    iters, status, nsteps = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
    pb = make_problem(kind, par, alg, dt, tol, dtmin, solver_maxiters, False)
    rc = lib(fast).orc_lyapunov(C.byref(pb), _dp(u0), n, float(t), float(ltol), float(lreltol), int(maxiters), _dp(lam), _dp(xf),
                                iters.ctypes.data_as(C.POINTER(C.c_int)), status.ctypes.data_as(C.POINTER(C.c_int)),
                                nsteps.ctypes.data_as(C.POINTER(C.c_long)))
    if rc:
        raise ValueError("bad interval")
    return lam, xf, iters, status, nsteps


def integrate_events(kind, par, u0, t_end, coef, offset=0.0, direction=0, max_events=1024, alg=ALG_DOP853, dt=0.0, tol=1e-12,
                     dtmin=None, maxiters=10 ** 9, backwards=False, fast=False):
    """surface-of-section crossings of coef.u - offset = 0 on [0, t_end] (docs/src/ClassicalDickeExamples.md:82-109):
    (ev_t[n][cap], ev_u[n][cap][nv], ev_dir[n][cap], ev_count[n], u_end[n][nv], status[n], attempts[n])"""
    nv = nvar(kind)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
    n, cap = u0.shape[0], int(max_events)
    cf = np.zeros(4)
    cf[:nv] = coef
    ev_t, ev_u = np.full((n, cap), np.nan), np.full((n, cap, nv), np.nan)
    ev_dir, cnt = np.zeros((n, cap), dtype=np.int32), np.empty(n, dtype=np.int32)
    u_end, status, nsteps = np.empty((n, nv)), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
    pb = make_problem(kind, par, alg, dt, tol, dtmin, maxiters, backwards)
    ip, lp = C.POINTER(C.c_int), C.POINTER(C.c_long)
    rc = lib(fast).orc_integrate_events(C.byref(pb), _dp(u0), n, float(t_end), _dp(cf), float(offset), int(direction), cap, _dp(ev_t),
                                        _dp(ev_u), ev_dir.ctypes.data_as(ip), cnt.ctypes.data_as(ip), _dp(u_end),
                                        status.ctypes.data_as(ip), nsteps.ctypes.data_as(lp))
    if rc:
        raise ValueError("bad t_end")
    return ev_t, ev_u, ev_dir, cnt, u_end, status, nsteps


def shell_quadrature(par, eps, Q, P, exprs, p_res=0.01, onlyqroot=0):
    """∫∫dqdpδϵ, src/EnergyShellProjections.jl:49-96, statement for statement (sequential sums): returns the list of
    values for `exprs` (expression strings over Q,q,P,p) or None where the reference returns `nonvalue`."""
    w0, w, g = par
    if Q ** 2 + P ** 2 > 4:
        return None
    A = (4 * g ** 2 * Q ** 2 * (1 - (Q ** 2 + P ** 2) / 4) - w * w0 * (P ** 2 + Q ** 2 - 2) + 2 * w * eps)
    if A < 0:
        return None
    pp = math.sqrt(A) / w - 1e-10
    if pp < 0:
        return None
    n = int(2 * math.floor(pp / p_res) + 1)
    val = None
    for sgn in (+1, -1):
        if onlyqroot not in (0, None) and onlyqroot != sgn:
            continue
        k = np.arange(1, n + 1)
        p = pp * np.cos(math.pi * (2 * k - 1) / (2 * n))
        r = 1 - (Q ** 2 + P ** 2) / 4
        disc = 4 * g ** 2 * Q ** 2 * r - p ** 2 * w ** 2 - w * w0 * (P ** 2 + Q ** 2 - 2) + 2 * w * eps   # src/ClassicalDicke.jl:224-228
        q = (-2 * g * Q * math.sqrt(r) + sgn * np.sqrt(disc)) / w                                        # :271-272
        X = np.column_stack([np.full(n, Q), q, np.full(n, P), p])
        F = [np.broadcast_to(eval_expr(e, ["Q", "q", "P", "p"], X), (n,)) for e in exprs]   # f at every node
        for kk in range(n):                                                                  # the reference's sequential sum
            mv = [(math.pi / (n * w)) * float(Fe[kk]) for Fe in F]
            val = mv if val is None else [a + b for a, b in zip(val, mv)]
    return val


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def uniforms(seed, idx, attempt=0):
    U = np.empty(4)
    lib().orc_uniforms(int(seed), int(idx), int(attempt), _dp(U))
    return U


def sample(kind, center, hbar, n, seed=0, offset=0, attempt=0):
    nv = 4 if kind == SAMPLER_HWXSU2 else 2
    s = make_sampler(kind, center, hbar, seed, offset)
    u = np.empty((n, nv))
    lib().orc_sample(C.byref(s), n, attempt, _dp(u))
    return u


def pdf(kind, center, hbar, u):
    nv = 4 if kind == SAMPLER_HWXSU2 else 2
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nv)
    s = make_sampler(kind, center, hbar)
    w = np.empty(u.shape[0])
    lib().orc_pdf_batch(C.byref(s), _dp(u), u.shape[0], _dp(w))
    return w


def scale_to_index(v, start, stop, length):
    """src/TWA.jl:53-64; returns 1-based indices, 0 where the reference returns NaN."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    idx = np.empty(v.size, dtype=np.int64)
    lib().orc_scale_to_index_batch(_dp(v), v.size, float(start), float(stop), int(length),
                                   idx.ctypes.data_as(C.POINTER(C.c_long)))
    return idx.reshape(v.shape)


# ------------------------------------------------------------------------------------------------
# Observables.  The reference turns an expression into a Julia lambda (src/TWA.jl:277-299); here an
# expression string over the system's varnames is walked with Python's `ast` and evaluated with
# numpy float64 scalar ops (every +,-,*,/,sqrt is one correctly rounded IEEE operation, no FMA).
# Integer powers follow the rule written in DESIGN.md: x^2 = x*x, x^3 = (x*x)*x, x^4 = (x*x)*(x*x),
# other integers by binary powering, non-integers by pow().
_FUNCS = {"sqrt": np.sqrt, "sin": np.sin, "cos": np.cos, "tan": np.tan, "exp": np.exp, "log": np.log,
          "abs": np.abs, "acos": np.arccos, "asin": np.arcsin, "atan": np.arctan}


def _ipow(x, n):
    if n == 0:
        return np.ones_like(x)
    if n < 0:
        return 1.0 / _ipow(x, -n)
    if n == 1:
        return x
    if n == 2:
        return x * x
    if n == 3:
        return (x * x) * x
    h = _ipow(x, n // 2)
    r = h * h
    return r * x if n % 2 else r


def eval_expr(expr: str, names, U):
    """Evaluate `expr` (operators + - * / ^, functions in _FUNCS, numbers, names) on U[..., nvar]."""
    U = np.asarray(U, dtype=np.float64)
    env = {n: U[..., i] for i, n in enumerate(names)}
    tree = ast.parse(expr.replace("^", "**"), mode="eval")

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant):
            return np.float64(node.value)
        if isinstance(node, ast.Name):
            if node.id == "pi":
                return np.float64(math.pi)
            return env[node.id]
        if isinstance(node, ast.UnaryOp):
            v = ev(node.operand)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.BinOp):
            if isinstance(node.op, ast.Pow):
                b = ev(node.left)
                e = node.right
                neg = False
                if isinstance(e, ast.UnaryOp) and isinstance(e.op, ast.USub):
                    neg, e = True, e.operand
                if isinstance(e, ast.Constant) and float(e.value) == int(e.value):
                    n = int(e.value)
                    return _ipow(b, -n if neg else n)
                return np.power(b, ev(node.right))
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, ast.Add):
                return a + b
            if isinstance(node.op, ast.Sub):
                return a - b
            if isinstance(node.op, ast.Mult):
                return a * b
            if isinstance(node.op, ast.Div):
                return a / b
        if isinstance(node, ast.Call):
            return _FUNCS[node.func.id](*[ev(a) for a in node.args])
        raise ValueError(f"unsupported syntax in {expr!r}")

    with np.errstate(all="ignore"):
        out = ev(tree)
    return np.broadcast_to(out, U.shape[:-1]).astype(np.float64)


class Weyl:
    """src/TWA.jl:789-884, as expression strings in the source's own operation order."""
    _cos = "((P^2+Q^2)/2-1)"
    _sin = f"sqrt(1-{_cos}^2)"
    _cph = "(Q/sqrt(P^2+Q^2))"
    _sph = "(-P/sqrt(P^2+Q^2))"

    @staticmethod
    def n(j):
        return f"({float(j)!r}*(p^2+q^2)-1)/2"

    @staticmethod
    def n2(j):
        return f"({Weyl.n(j)})^2-1/4"

    @staticmethod
    def Jz(j):
        return f"{math.sqrt(j * (j + 1))!r}*{Weyl._cos}"

    @staticmethod
    def Jx(j):
        return f"{math.sqrt(j * (j + 1))!r}*{Weyl._sin}*{Weyl._cph}"

    @staticmethod
    def Jy(j):
        return f"{math.sqrt(j * (j + 1))!r}*{Weyl._sin}*{Weyl._sph}"

    @staticmethod
    def _sq(j, inner):
        bj = math.sqrt(j * (j + 1) * (2 * j - 1) * (2 * j + 3))
        return f"{(j * (j + 1)) / 3!r}+{bj / 2!r}*({inner}^2-1/3)"

    @staticmethod
    def Jz2(j):
        return Weyl._sq(j, Weyl._cos)

    @staticmethod
    def Jx2(j):
        return Weyl._sq(j, f"({Weyl._sin}*{Weyl._cph})")

    @staticmethod
    def Jy2(j):
        return Weyl._sq(j, f"({Weyl._sin}*{Weyl._sph})")


# ------------------------------------------------------------------------------------------------
# Reducers on dense trajectories traj[n][nt][nvar] (NaN rows = the trajectory contributed nothing at
# that time: failure or maxiters).
def _valid(traj):
    return np.isfinite(traj).all(axis=2)


def average_sums(values, valid):
    """a14 src/TWA.jl:355-360: per-time sums (not yet divided by N) in trajectory order."""
    v = np.where(valid, values, 0.0)
    out = np.zeros(v.shape[1])
    for i in range(v.shape[0]):  # sequential accumulation order of the reference's worker loop
        out += v[i]
    return out


def hist_time(values, valid, start, stop, length):
    """a18, x=:t or y=:t modes (src/TWA.jl:456-468): counts[nt][len]."""
    n, nt = values.shape
    idx = scale_to_index(values, start, stop, length)
    counts = np.zeros((nt, length), dtype=np.uint64)
    for k in range(nt):
        ok = valid[:, k] & (idx[:, k] > 0)
        np.add.at(counts[k], idx[ok, k] - 1, 1)
    return counts


def hist_2d(xv, yv, valid, xr, yr, animate):
    """a18, neither axis is :t (src/TWA.jl:470-489): counts[iy][ix] over all times, or per time."""
    n, nt = xv.shape
    ix = scale_to_index(xv, *xr)
    iy = scale_to_index(yv, *yr)
    ok = valid & (ix > 0) & (iy > 0)
    if animate:
        counts = np.zeros((nt, yr[2], xr[2]), dtype=np.uint64)
        for k in range(nt):
            np.add.at(counts[k], (iy[ok[:, k], k] - 1, ix[ok[:, k], k] - 1), 1)
    else:
        counts = np.zeros((yr[2], xr[2]), dtype=np.uint64)
        np.add.at(counts, (iy[ok] - 1, ix[ok] - 1), 1)
    return counts


def ensemble(kind, par, sampler, N, ts, values, hists=(), alg=ALG_RK4, dt=2.0 ** -7, tol=1e-12, dtmin=None,
             maxiters=100000, backwards=False, host_u0=None, workers=1, tolerate_errors=True, fast=False, controller=None, output=None):
    """The C driver (src/TWA.jl:100-224 restated).  values: list of (kind, a, c); hists: list of
    (value_index, start, stop, len).  Returns a dict of raw sums / counts / counters."""
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    nt = len(ts)
    pb = make_problem(kind, par, alg, dt, tol, dtmin, maxiters, backwards, controller, output)
    vals = (Value * len(values))(*[Value(k, a, c) for (k, a, c) in values])
    hs = (Hist * max(1, len(hists)))(*[Hist(v, a, b, n) for (v, a, b, n) in hists])
    sums = np.zeros((nt, len(values)))
    counts = [np.zeros((nt, h[3]), dtype=np.uint64) for h in hists]
    cptr = (C.POINTER(C.c_uint64) * max(1, len(hists)))(*[c.ctypes.data_as(C.POINTER(C.c_uint64)) for c in counts])
    nvalid = np.zeros(nt, dtype=np.uint64)
    a, r, f = C.c_long(), C.c_long(), C.c_long()
    u0p = None
    if host_u0 is not None:
        host_u0 = np.ascontiguousarray(host_u0, dtype=np.float64)
        u0p = _dp(host_u0)
    rc = lib(fast).orc_ensemble(C.byref(pb), C.byref(sampler) if sampler is not None else None, u0p, int(N),
                                _dp(ts), nt, vals, len(values), hs, len(hists), int(workers),
                                int(tolerate_errors), _dp(sums), cptr,
                                nvalid.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(a), C.byref(r), C.byref(f))
    if rc not in (0,):
        raise RuntimeError(f"orc_ensemble failed ({rc})")
    return {"sums": sums, "counts": counts, "n_valid": nvalid, "steps_accepted": a.value,
            "steps_rejected": r.value, "n_failed": f.value}


def max_threads():
    return int(lib().orc_max_threads())
