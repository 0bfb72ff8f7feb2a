"""ctypes binding of libdicketwa.so (include/dicketwa.h) -- the Python twin of the Julia `ccall` shim.

Loads the in-tree shared object `dickemodel.jl_b200/csrc/libdicketwa.so` and fails loudly when it is
missing or when no CUDA device exists: there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIBDICKETWA") or os.path.join(HERE, "csrc", "libdicketwa.so")   # (the override is for A/B builds)

# ---- constants (include/dicketwa.h) ----
OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_EXPR, ERR_NCCL, ERR_TOO_MANY_FAILURES, ERR_BOUNDS, ERR_CAPACITY, ERR_JIT = range(-1, -10, -1)
TRAJ_OK, TRAJ_DOMAIN, TRAJ_MAXITERS, TRAJ_BAD_IC = 0, 1, 2, 3
LYAP_MAXITERS = 5
SYS_DICKE, SYS_LMG = 0, 1
ALG_RK4, ALG_RK8, ALG_DP5, ALG_DOP853, ALG_TSIT5 = 0, 1, 2, 3, 4
CTRL_ODE, CTRL_SCIPY = 0, 1
CHART_QP, CHART_BLOCH = 0, 1
OUTPUT_STEP, OUTPUT_HERMITE = 0, 1
SAMPLER_HOST_ARRAY, SAMPLER_COHERENT_HW, SAMPLER_COHERENT_SU2, SAMPLER_COHERENT_HWXSU2, SAMPLER_DEVICE_ARRAY = 0, 1, 2, 3, 4
RED_AVERAGE, RED_HIST, RED_SURVIVAL = 0, 1, 2
AXIS_EXPR, AXIS_TIME = 0, 1

EXPORTS = ["dtwa_last_error", "dtwa_version", "dtwa_create", "dtwa_destroy", "dtwa_device_info", "dtwa_set_launch",
           "dtwa_hamiltonian", "dtwa_integrate", "dtwa_sample", "dtwa_pdf", "dtwa_eval_expr", "dtwa_scale_to_index",
           "dtwa_ensemble", "dtwa_ensemble_launch", "dtwa_plan_arenas", "dtwa_plan_fetch", "dtwa_plan_free",
           "dtwa_fp64_peak", "dtwa_fp64_peak_regs", "dtwa_fp64_chains", "dtwa_fp64_mix", "dtwa_set_jit", "dtwa_jit_compile", "dtwa_integrate_fm", "dtwa_lyapunov", "dtwa_integrate_events",
           "dtwa_integrate_events_packed", "dtwa_integrate_events_packed_fetch", "dtwa_shell_quadrature", "dtwa_jit_quiesce"]


class Event(C.Structure):
    """dtwa_event: the section coef . u - offset = 0"""
    _fields_ = [("coef", C.c_double * 4), ("offset", C.c_double), ("direction", C.c_int32), ("pad", C.c_int32)]


class DtwaError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference raises Julia error(...))."""

    def __init__(self, code, msg):
        super().__init__(f"libdicketwa error {code}: {msg}")
        self.code = code


class System(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("params", C.c_double * 4)]


class Solver(C.Structure):
    _fields_ = [("alg", C.c_int32), ("backwards", C.c_int32), ("dt", C.c_double), ("tol", C.c_double),
                ("dtmin", C.c_double), ("maxiters", C.c_int64), ("controller", C.c_int32), ("chart", C.c_int32),
                ("beta1", C.c_double), ("beta2", C.c_double), ("qmin", C.c_double), ("qmax", C.c_double),
                ("gamma", C.c_double), ("qoldinit", C.c_double), ("output", C.c_int32), ("reserved", C.c_int32)]


class Sampler(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("center", C.c_double * 4), ("hbar", C.c_double),
                ("seed", C.c_uint64), ("offset", C.c_uint64), ("u0", C.c_void_p)]


class Range(C.Structure):
    _fields_ = [("start", C.c_double), ("stop", C.c_double), ("len", C.c_int64)]


class Reducer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nexpr", C.c_int32), ("exprs", C.POINTER(C.c_char_p)),
                ("x_axis", C.c_int32), ("y_axis", C.c_int32), ("x_expr", C.c_char_p), ("y_expr", C.c_char_p),
                ("xs", Range), ("ys", Range), ("animate", C.c_int32), ("reserved", C.c_int32)]


class ReducerOut(C.Structure):
    _fields_ = [("sums", C.POINTER(C.c_double)), ("sums_len", C.c_int64),
                ("counts", C.POINTER(C.c_uint64)), ("counts_len", C.c_int64)]


class Result(C.Structure):
    _fields_ = [("red", C.POINTER(ReducerOut)), ("n_valid", C.POINTER(C.c_uint64)),
                ("n_failed", C.c_uint64), ("n_unfinished", C.c_uint64), ("steps_accepted", C.c_uint64),
                ("steps_rejected", C.c_uint64), ("lane_steps", C.c_uint64), ("resample_rounds", C.c_int32),
                ("launches", C.c_int32), ("jit", C.c_int32), ("reserved", C.c_int32), ("kernel_ms", C.c_double),
                ("total_ms", C.c_double)]


_lib = None
_lock = threading.Lock()


def load():
    """dlopen the in-tree library and declare every prototype of include/dicketwa.h."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a).  This package has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        L.dtwa_last_error.restype = C.c_char_p
        L.dtwa_version.restype = C.c_int
        L.dtwa_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, vp]
        L.dtwa_destroy.argtypes = [vp]
        L.dtwa_device_info.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.dtwa_set_launch.argtypes = [vp, C.c_int, C.c_int]
        L.dtwa_hamiltonian.argtypes = [vp, C.POINTER(System), dp, C.c_int64, dp]
        L.dtwa_integrate.argtypes = [vp, C.POINTER(System), C.POINTER(Solver), dp, C.c_int64, dp, C.c_int32, dp,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.dtwa_sample.argtypes = [vp, C.POINTER(Sampler), C.c_int64, C.c_uint32, dp]
        L.dtwa_pdf.argtypes = [vp, C.POINTER(Sampler), dp, C.c_int64, dp]
        L.dtwa_eval_expr.argtypes = [vp, C.POINTER(System), C.c_char_p, dp, C.c_int64, dp]
        L.dtwa_scale_to_index.argtypes = [vp, dp, C.c_int64, C.POINTER(Range), C.POINTER(C.c_int64)]
        L.dtwa_ensemble.argtypes = [vp, C.POINTER(System), C.POINTER(Solver), C.POINTER(Sampler), C.c_int64, dp,
                                    C.c_int32, C.POINTER(Reducer), C.c_int32, C.c_int32, C.POINTER(Result)]
        L.dtwa_ensemble_launch.argtypes = [vp, C.POINTER(System), C.POINTER(Solver), C.POINTER(Sampler), C.c_int64, dp,
                                           C.c_int32, C.POINTER(Reducer), C.c_int32, C.c_int32, C.POINTER(vp)]
        L.dtwa_plan_arenas.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(vp), C.POINTER(C.c_int64)]
        L.dtwa_plan_fetch.argtypes = [vp, C.POINTER(Result)]
        L.dtwa_plan_free.argtypes = [vp]
        L.dtwa_fp64_peak.argtypes = [vp, dp, dp]
        L.dtwa_fp64_peak_regs.argtypes = [vp, dp, dp]
        L.dtwa_fp64_chains.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp]
        L.dtwa_fp64_mix.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, dp]
        L.dtwa_jit_quiesce.argtypes = []
        import atexit
        atexit.register(L.dtwa_jit_quiesce)   # no thread may be inside libnvrtc when the process tears it down
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.dtwa_integrate_fm.argtypes = [vp, C.POINTER(System), C.POINTER(Solver), dp, dp, C.c_int64, dp, C.c_int32, dp, dp, i32p, i64p]
        L.dtwa_integrate_events.argtypes = [vp, C.POINTER(System), C.POINTER(Solver), dp, C.c_int64, C.c_double, C.POINTER(Event),
                                            C.c_int32, dp, dp, i32p, i32p, dp, i32p, i64p, dp]
        L.dtwa_integrate_events_packed.argtypes = [vp, C.POINTER(System), C.POINTER(Solver), dp, C.c_int64, C.c_double, C.POINTER(Event),
                                                   C.c_int32, i64p, i32p, dp, i32p, i64p, dp]
        L.dtwa_integrate_events_packed_fetch.argtypes = [vp, C.c_int64, dp, dp, i32p]
        L.dtwa_shell_quadrature.argtypes = [vp, C.POINTER(System), C.c_double, dp, C.c_int64, C.POINTER(C.c_char_p), C.c_int32, C.c_double,
                                            C.c_int32, dp, i32p, i32p, dp]
        L.dtwa_lyapunov.argtypes = [vp, C.POINTER(System), C.POINTER(Solver), dp, C.c_int64, C.c_double, C.c_double, C.c_double,
                                    C.c_int32, dp, dp, i32p, i32p, i64p, dp]
        L.dtwa_set_jit.argtypes = [vp, C.c_int]
        L.dtwa_jit_compile.argtypes = [C.POINTER(System), C.POINTER(Solver), C.POINTER(C.c_char_p), C.c_int32, C.c_char_p,
                                       C.c_int64, C.POINTER(C.c_int64)]
        for name in EXPORTS:
            getattr(L, name).restype = C.c_char_p if name == "dtwa_last_error" else C.c_int
        _lib = L
        return L


def _check(rc):
    if rc != OK:
        raise DtwaError(rc, load().dtwa_last_error().decode("utf-8", "replace"))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """Owns a dtwa_ctx (device memory, stream).  One per process and device set."""

    def __init__(self, devices=None, stream=None):
        L = load()
        h = C.c_void_p()
        if devices is None:
            devices = [0]
        ids = (C.c_int * len(devices))(*devices)
        # stream: None -> the library creates a private non-blocking stream; an int is a cudaStream_t handle,
        # where 0 (the legacy default stream, what torch.cuda.current_stream().cuda_stream returns by default)
        # is passed as cudaStreamLegacy (0x1) because NULL means "private stream" in the C ABI
        sp = None if stream is None else C.c_void_p(int(stream) if int(stream) != 0 else 1)
        _check(L.dtwa_create(C.byref(h), ids, len(devices), sp))
        self._h = h
        self.devices = list(devices)
        self._packed_lock = threading.Lock()

    def close(self):
        if getattr(self, "_h", None):
            load().dtwa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- info / tuning ----
    def device_info(self, which=0):
        name = C.create_string_buffer(256)
        sm, clk, cc = C.c_int(), C.c_int(), C.c_int()
        _check(load().dtwa_device_info(self._h, which, name, C.byref(sm), C.byref(clk), C.byref(cc)))
        return {"name": name.value.decode(), "sm_count": sm.value, "clock_khz": clk.value, "cc": cc.value}

    def set_launch(self, block_threads=0, ctas_per_sm=0):
        _check(load().dtwa_set_launch(self._h, block_threads, ctas_per_sm))

    def set_jit(self, mode):
        """0: interpret the reducer program, 1: always specialise with NVRTC, 2: auto (default)"""
        _check(load().dtwa_set_jit(self._h, int(mode)))

    def fp64_peak(self):
        tf, ms = C.c_double(), C.c_double()
        _check(load().dtwa_fp64_peak(self._h, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def fp64_chains(self, chains, warps_per_sm, nreg=1):
        tf, ms = C.c_double(), C.c_double()
        _check(load().dtwa_fp64_chains(self._h, int(chains), int(warps_per_sm), int(nreg), C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def fp64_mix(self, alu_per_8, kind, warps_per_sm):
        tf, ms = C.c_double(), C.c_double()
        _check(load().dtwa_fp64_mix(self._h, int(alu_per_8), int(kind), int(warps_per_sm), C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def fp64_peak_regs(self):
        tf, ms = C.c_double(), C.c_double()
        _check(load().dtwa_fp64_peak_regs(self._h, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    # ---- batch helpers ----
    def hamiltonian(self, sys: System, u):
        nv = 4 if sys.kind == SYS_DICKE else 2
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nv)
        h = np.empty(u.shape[0])
        _check(load().dtwa_hamiltonian(self._h, C.byref(sys), _dp(u), u.shape[0], _dp(h)))
        return h

    def integrate(self, sys: System, sol: Solver, u0, ts):
        nv = 4 if sys.kind == SYS_DICKE else 2
        u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        n, nt = u0.shape[0], ts.shape[0]
        out = np.empty((n, nt, nv))
        status = np.empty(n, dtype=np.int32)
        nsteps = np.empty(n, dtype=np.int64)
        _check(load().dtwa_integrate(self._h, C.byref(sys), C.byref(sol), _dp(u0), n, _dp(ts), nt, _dp(out),
                                     status.ctypes.data_as(C.POINTER(C.c_int32)),
                                     nsteps.ctypes.data_as(C.POINTER(C.c_int64))))
        return out, status, nsteps

    def integrate_fm(self, sys: System, sol: Solver, u0, ts, fm0=None):
        """(x[n][nt][nv], Phi[n][nt][nv][nv] (Phi[..., r, c]), status, nsteps); fm0[n][nv][nv] or None = identity"""
        nv = 4 if sys.kind == SYS_DICKE else 2
        u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        n, nt = u0.shape[0], ts.shape[0]
        f0 = None
        if fm0 is not None:   # row/column indexed -> column-major flat (Julia's layout)
            f0 = np.ascontiguousarray(np.swapaxes(np.asarray(fm0, dtype=np.float64).reshape(n, nv, nv), 1, 2))
        out_u, out_fm = np.empty((n, nt, nv)), np.empty((n, nt, nv * nv))
        status, nsteps = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
        _check(load().dtwa_integrate_fm(self._h, C.byref(sys), C.byref(sol), _dp(u0), _dp(f0) if f0 is not None else None, n,
                                        _dp(ts), nt, _dp(out_u), _dp(out_fm), status.ctypes.data_as(C.POINTER(C.c_int32)),
                                        nsteps.ctypes.data_as(C.POINTER(C.c_int64))))
        return out_u, np.swapaxes(out_fm.reshape(n, nt, nv, nv), 2, 3), status, nsteps

    def integrate_events(self, sys: System, sol: Solver, u0, t_end, coef, offset=0.0, direction=0, max_events=1024):
        """(ev_t[n][cap], ev_u[n][cap][nv], ev_dir[n][cap], ev_count[n], u_end[n][nv], status[n], nsteps[n], kernel_ms)"""
        nv = 4 if sys.kind == SYS_DICKE else 2
        u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
        n, cap = u0.shape[0], int(max_events)
        ev = Event()
        for i, c in enumerate(np.asarray(coef, dtype=float).ravel()[:nv]):
            ev.coef[i] = float(c)
        ev.offset, ev.direction = float(offset), int(direction)
        ev_t, ev_u = np.empty((n, cap)), np.empty((n, cap, nv))      # (slots beyond ev_count are set to NaN / 0 below)
        ev_dir, cnt = np.empty((n, cap), dtype=np.int32), np.empty(n, dtype=np.int32)
        u_end, status, nsteps = np.empty((n, nv)), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
        ms = C.c_double()
        i32p = C.POINTER(C.c_int32)
        _check(load().dtwa_integrate_events(self._h, C.byref(sys), C.byref(sol), _dp(u0), n, float(t_end), C.byref(ev), cap, _dp(ev_t),
                                            _dp(ev_u), ev_dir.ctypes.data_as(i32p), cnt.ctypes.data_as(i32p), _dp(u_end),
                                            status.ctypes.data_as(i32p), nsteps.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(ms)))
        valid = np.arange(cap)[None, :] < np.minimum(cnt, cap)[:, None]   # slots beyond ev_count hold scratch
        ev_t[~valid] = np.nan
        ev_u[~valid] = np.nan
        ev_dir[~valid] = 0
        return ev_t, ev_u, ev_dir, cnt, u_end, status, nsteps, ms.value

    def integrate_events_packed(self, sys: System, sol: Solver, u0, t_end, coef, offset=0.0, direction=0, max_events=1024):
        """CSR output: (offsets[n+1], ev_t[total], ev_u[total][nv], ev_dir[total], ev_count[n], u_end[n][nv], status[n],
        nsteps[n], kernel_ms); row i owns [offsets[i], offsets[i+1])"""
        nv = 4 if sys.kind == SYS_DICKE else 2
        u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
        n, cap = u0.shape[0], int(max_events)
        ev = Event()
        for i, c in enumerate(np.asarray(coef, dtype=float).ravel()[:nv]):
            ev.coef[i] = float(c)
        ev.offset, ev.direction = float(offset), int(direction)
        offsets, cnt = np.zeros(n + 1, dtype=np.int64), np.empty(n, dtype=np.int32)
        u_end, status, nsteps = np.empty((n, nv)), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
        ms = C.c_double()
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        with self._packed_lock:      # the packed arrays wait on the device between the two calls
            _check(load().dtwa_integrate_events_packed(self._h, C.byref(sys), C.byref(sol), _dp(u0), n, float(t_end), C.byref(ev), cap,
                                                       offsets.ctypes.data_as(i64p), cnt.ctypes.data_as(i32p), _dp(u_end),
                                                       status.ctypes.data_as(i32p), nsteps.ctypes.data_as(i64p), C.byref(ms)))
            total = int(offsets[n])
            ev_t, ev_u, ev_dir = np.empty(total), np.empty((total, nv)), np.empty(total, dtype=np.int32)
            _check(load().dtwa_integrate_events_packed_fetch(self._h, total, _dp(ev_t), _dp(ev_u), ev_dir.ctypes.data_as(i32p)))
        return offsets, ev_t, ev_u, ev_dir, cnt, u_end, status, nsteps, ms.value

    def shell_quadrature(self, sys: System, eps, QP, exprs, p_res=0.01, onlyqroot=0):
        """(out[ncell][nexpr] (NaN = outside the shell), valid[ncell], nodes[ncell], kernel_ms)"""
        QP = np.ascontiguousarray(QP, dtype=np.float64).reshape(-1, 2)
        n = QP.shape[0]
        ex = [e.encode("utf-8") for e in exprs]
        arr = (C.c_char_p * len(ex))(*ex)
        out = np.empty((n, len(ex)))
        valid, nodes = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        ms = C.c_double()
        i32p = C.POINTER(C.c_int32)
        _check(load().dtwa_shell_quadrature(self._h, C.byref(sys), float(eps), _dp(QP), n, arr, len(ex), float(p_res), int(onlyqroot),
                                            _dp(out), valid.ctypes.data_as(i32p), nodes.ctypes.data_as(i32p), C.byref(ms)))
        return out, valid, nodes, ms.value

    def lyapunov(self, sys: System, sol: Solver, u0, t, ltol, lreltol, maxiters):
        """(lambda[n][nv], x_final[n][nv], iters[n], status[n], nsteps[n], kernel_ms)"""
        nv = 4 if sys.kind == SYS_DICKE else 2
        u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
        n = u0.shape[0]
        lam, xf = np.empty((n, nv)), np.empty((n, nv))
        iters, status, nsteps = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
        ms = C.c_double()
        _check(load().dtwa_lyapunov(self._h, C.byref(sys), C.byref(sol), _dp(u0), n, float(t), float(ltol), float(lreltol),
                                    int(maxiters), _dp(lam), _dp(xf), iters.ctypes.data_as(C.POINTER(C.c_int32)),
                                    status.ctypes.data_as(C.POINTER(C.c_int32)), nsteps.ctypes.data_as(C.POINTER(C.c_int64)),
                                    C.byref(ms)))
        return lam, xf, iters, status, nsteps, ms.value

    def sample(self, smp: Sampler, n, attempt=0):
        nv = 4 if smp.kind == SAMPLER_COHERENT_HWXSU2 else 2
        u = np.empty((n, nv))
        _check(load().dtwa_sample(self._h, C.byref(smp), n, attempt, _dp(u)))
        return u

    def pdf(self, smp: Sampler, u):
        nv = 4 if smp.kind == SAMPLER_COHERENT_HWXSU2 else 2
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nv)
        w = np.empty(u.shape[0])
        _check(load().dtwa_pdf(self._h, C.byref(smp), _dp(u), u.shape[0], _dp(w)))
        return w

    def eval_expr(self, sys: System, expr: str, u):
        nv = 4 if sys.kind == SYS_DICKE else 2
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nv)
        v = np.empty(u.shape[0])
        _check(load().dtwa_eval_expr(self._h, C.byref(sys), expr.encode("utf-8"), _dp(u), u.shape[0], _dp(v)))
        return v

    def scale_to_index(self, v, start, stop, length):
        v = np.ascontiguousarray(v, dtype=np.float64)
        idx = np.empty(v.size, dtype=np.int64)
        r = Range(float(start), float(stop), int(length))
        _check(load().dtwa_scale_to_index(self._h, _dp(v), v.size, C.byref(r), idx.ctypes.data_as(C.POINTER(C.c_int64))))
        return idx.reshape(v.shape)

    # ---- the ensemble ----
    def ensemble(self, sys: System, sol: Solver, smp: Sampler, N, ts, reducers, tolerate_errors=True, allreduce=None):
        """reducers: list of dicts (see TWA._spec_*).  Returns (list of raw arrays, info dict).

        allreduce: optional callable (d_f64_ptr, n_f64, d_u64_ptr, n_u64) that sums the two device
        arenas over the ranks of a one-process-per-GPU job (torch.distributed NCCL in bench.py)."""
        L = load()
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        nt = ts.shape[0]
        keep = []  # keep ctypes objects alive
        reds = (Reducer * len(reducers))()
        outs = (ReducerOut * len(reducers))()
        arrays = []
        for i, r in enumerate(reducers):
            R = reds[i]
            R.kind = r["kind"]
            if r["kind"] == RED_AVERAGE:
                ex = [e.encode("utf-8") for e in r["exprs"]]
                arr = (C.c_char_p * len(ex))(*ex)
                keep.append((ex, arr))
                R.nexpr = len(ex)
                R.exprs = arr
                a = np.zeros((nt, len(ex)))
                outs[i].sums = _dp(a)
                outs[i].sums_len = a.size
            elif r["kind"] == RED_SURVIVAL:
                a = np.zeros((nt, 1))
                outs[i].sums = _dp(a)
                outs[i].sums_len = a.size
            else:
                R.x_axis, R.y_axis = r["x_axis"], r["y_axis"]
                if r.get("x_expr") is not None:
                    R.x_expr = r["x_expr"].encode("utf-8")
                if r.get("y_expr") is not None:
                    R.y_expr = r["y_expr"].encode("utf-8")
                if r.get("xs") is not None:
                    R.xs = Range(*r["xs"])
                if r.get("ys") is not None:
                    R.ys = Range(*r["ys"])
                R.animate = int(bool(r.get("animate", False)))
                if R.x_axis == AXIS_TIME:
                    shape = (nt, int(R.ys.len))
                elif R.y_axis == AXIS_TIME:
                    shape = (nt, int(R.xs.len))
                elif R.animate:
                    shape = (nt, int(R.ys.len), int(R.xs.len))
                else:
                    shape = (int(R.ys.len), int(R.xs.len))
                a = np.zeros(shape, dtype=np.uint64)
                outs[i].counts = a.ctypes.data_as(C.POINTER(C.c_uint64))
                outs[i].counts_len = a.size
            arrays.append(a)
        n_valid = np.zeros(nt, dtype=np.uint64)
        res = Result()
        res.red = outs
        res.n_valid = n_valid.ctypes.data_as(C.POINTER(C.c_uint64))
        if allreduce is None:
            _check(L.dtwa_ensemble(self._h, C.byref(sys), C.byref(sol), C.byref(smp), int(N), _dp(ts), nt, reds,
                                   len(reducers), int(bool(tolerate_errors)), C.byref(res)))
        else:
            plan = C.c_void_p()
            _check(L.dtwa_ensemble_launch(self._h, C.byref(sys), C.byref(sol), C.byref(smp), int(N), _dp(ts), nt, reds,
                                          len(reducers), int(bool(tolerate_errors)), C.byref(plan)))
            try:
                pf, pu = C.c_void_p(), C.c_void_p()
                nf, nu = C.c_int64(), C.c_int64()
                _check(L.dtwa_plan_arenas(plan, C.byref(pf), C.byref(nf), C.byref(pu), C.byref(nu)))
                allreduce(pf.value, nf.value, pu.value, nu.value)
                _check(L.dtwa_plan_fetch(plan, C.byref(res)))
            finally:
                L.dtwa_plan_free(plan)
        info = {"n_valid": n_valid, "n_failed": int(res.n_failed), "n_unfinished": int(res.n_unfinished),
                "steps_accepted": int(res.steps_accepted), "steps_rejected": int(res.steps_rejected),
                "lane_steps": int(res.lane_steps), "resample_rounds": int(res.resample_rounds),
                "launches": int(res.launches), "jit": int(res.jit), "kernel_ms": float(res.kernel_ms), "total_ms": float(res.total_ms)}
        return arrays, info


def jit_compile(sys: System, sol: Solver, exprs):
    """NVRTC compile check of the specialised ensemble kernel (no GPU needed): (source, cubin bytes)"""
    ex = [e.encode("utf-8") for e in exprs]
    arr = (C.c_char_p * len(ex))(*ex)
    buf = C.create_string_buffer(1 << 16)
    nb = C.c_int64()
    _check(load().dtwa_jit_compile(C.byref(sys), C.byref(sol), arr, len(ex), buf, len(buf), C.byref(nb)))
    return buf.value.decode(), nb.value


_default_ctx = None
_dist = {"rank": 0, "world": 1, "allreduce": None}


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context([int(os.environ.get("DTWA_DEVICE", "0"))])
        if "DTWA_JIT" in os.environ:
            _default_ctx.set_jit(int(os.environ["DTWA_JIT"]))
    return _default_ctx


_device_ctxs = {}


def device_context(dev: int) -> Context:
    """one cached single-device context per GPU (for batches split over several GPUs, see split_over_devices)"""
    dev = int(dev)
    if dev not in _device_ctxs:
        _device_ctxs[dev] = Context([dev])
    return _device_ctxs[dev]


def run_on_devices(devices, calls):
    """calls[k](context of devices[k]) for all k at once (host threads; ctypes releases the GIL inside the C library)"""
    import concurrent.futures as cf
    ctxs = [device_context(d) for d in devices]      # (created serially: context creation is not re-entrant)
    with cf.ThreadPoolExecutor(max_workers=len(calls)) as ex:
        return [f.result() for f in [ex.submit(call, c) for call, c in zip(calls, ctxs)]]


def split_over_devices(devices, n, call):
    """Independent batch items (initial conditions, (Q,P) cells) have no exchange step: GPU k of `devices` takes the items
    k, k+ndev, k+2 ndev, ... (interleaved, so that regions of expensive items spread evenly) and `call(ctx, items)` runs for
    all GPUs at once (ctypes releases the GIL while the C library runs).  `call` returns a tuple of per-item arrays
    (first axis = items) and scalars; arrays are scattered back into item order, scalars (times) reduced with max."""
    devices = list(devices)[:max(1, n)]
    items = [slice(k, n, len(devices)) for k in range(len(devices))]
    parts = run_on_devices(devices, [lambda c, it=it: call(c, it) for it in items])
    merged = []
    for j, first in enumerate(parts[0]):
        if isinstance(first, np.ndarray):
            full = np.empty((n,) + first.shape[1:], dtype=first.dtype)
            for it, p in zip(items, parts):
                full[it] = p[j]
            merged.append(full)
        else:
            merged.append(max(p[j] for p in parts))
    return tuple(merged)


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


def set_distributed(rank=0, world=1, allreduce=None):
    """One-process-per-GPU mode: monte_carlo_integrate gives rank r the trajectory block r of
    `world` equal blocks (Philox counters are global indices, so the sample set is the same for any
    world size) and calls `allreduce` on the device arenas before results are fetched."""
    _dist.update(rank=int(rank), world=int(world), allreduce=allreduce)


def distributed():
    return dict(_dist)
