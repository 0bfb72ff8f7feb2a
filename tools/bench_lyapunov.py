#!/usr/bin/env python
"""Lyapunov map throughput: the docs' map (docs/src/ClassicalDickeExamples.md:133-236,
w=0.8, g=0.8, w0=1, eps=-1.35) on an n x n grid of (Q, P) in one dtwa_lyapunov launch, next to the CPU
oracle on a bounded sample of the same points.  One JSON line.

usage: python tools/bench_lyapunov.py [--n 256] [--tol 1e-8] [--alg DP8] [--cpu-points 256]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--tol", type=float, default=1e-8)
ap.add_argument("--alg", default="DP8")
ap.add_argument("--dt", type=float, default=2.0 ** -6)
ap.add_argument("--cpu-points", type=int, default=256)
ap.add_argument("--devices", default="", help="comma-separated GPU ids: split the map over several GPUs (default: one)")
a = ap.parse_args()
pkg = graft.load_package()
CD, CS = pkg.ClassicalDicke, pkg.ClassicalSystems
system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
eps = -1.35
Qs = np.linspace(CD.minimum_nonnegative_Q_for_ϵ(system, eps), CD.maximum_Q_for_ϵ(system, eps), a.n)
Ps = np.linspace(0, CD.maximum_P_for_ϵ(system, eps), a.n)
pts = np.array([CD.Point(system, Q=Q, P=P, p=0, ϵ=eps) for Q in Qs for P in Ps if CD.minimum_ϵ_for(system, P=P, p=0, Q=Q) <= eps])
kw = dict(integrator_alg=a.alg, tol=a.tol, λtol=5e-5) if a.alg not in ("RK4", "RK8") else dict(integrator_alg=a.alg, dt=a.dt, λtol=5e-5)
if a.devices:
    kw["devices"] = [int(d) for d in a.devices.split(",")]
CS.lyapunov_spectrum_batch(system, pts[:1024 * max(1, len(kw.get("devices", [0])))], 100, **kw)   # warm-up
t0 = time.perf_counter()
lam, status = CS.lyapunov_spectrum_batch(system, pts, 100, **kw)
wall = time.perf_counter() - t0
info = CS.last_lyapunov_info
steps = int(info["nsteps"].sum())
out = {"what": "Lyapunov map, variational system (4 + 16 variables), one launch", "grid": a.n, "devices": kw.get("devices", [0]), "points": int(len(pts)), "alg": a.alg,
       "tol": a.tol, "variational_steps": steps, "kernel_ms": info["kernel_ms"], "wall_s": wall,
       "steps_per_s_kernel": steps / (info["kernel_ms"] * 1e-3), "points_per_s": len(pts) / wall,
       "converged": int((status == 0).sum()), "max_iterations_reached": int((status == 5).sum()),
       "mean_rounds": float(info["iterations"].mean()), "lambda_max_quantiles": [float(x) for x in np.quantile(lam.max(axis=1), [0.1, 0.5, 0.9])]}
if a.cpu_points and a.alg not in ("RK4", "RK8"):
    from oracle import oracle as O
    sel = np.random.default_rng(0).choice(len(pts), size=min(a.cpu_points, len(pts)), replace=False)
    t0 = time.perf_counter()
    lo, xo, ito, sto, nso = O.lyapunov(O.SYS_DICKE, [1.0, 0.8, 0.8], pts[sel], t=100, ltol=5e-5, lreltol=1e-3, maxiters=100, tol=a.tol,
                                       alg=O.ALG_DOP853 if a.alg == "DP8" else O.ALG_DP5, fast=True)
    cw = time.perf_counter() - t0
    same = (ito == info["iterations"][sel]) & (sto == status[sel])
    out["cpu_baseline"] = {"kind": "port", "cores": O.lib(True).orc_max_threads(), "points": int(len(sel)), "seconds": cw,
                           "steps_per_s": float(nso.sum() / cw), "points_per_s": len(sel) / cw,
                           "same_rounds_fraction": float(same.mean()),
                           "max_abs_diff_lambda_max_where_same": float(np.abs(lam[sel][same].max(axis=1) - lo[same].max(axis=1)).max())}
print(json.dumps(out))
