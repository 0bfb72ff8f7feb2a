#!/usr/bin/env python
"""Throughput of the SURVEY 8f rank 3/4 kernels next to the CPU oracle on a bounded sample (one JSON line each).

  sections : docs' Poincare surface (docs/src/ClassicalDickeExamples.md:82-109: w=0.8, g=0.8, w0=1, eps=-1.35, section p=0),
             n initial conditions on the shell, t = 1000, DOP853 tol=1e-10
  shell    : matrix_QP∫∫dqdpδϵ of f = [1, q^2+p^2] at res (src/EnergyShellProjections.jl:159-199)
usage: python tools/bench_widen.py [--n 100000] [--res 0.01]
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
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--res", type=float, default=0.01)
a = ap.parse_args()
pkg = graft.load_package()
CD, CS, ESP = pkg.ClassicalDicke, pkg.ClassicalSystems, pkg.EnergyShellProjections
from oracle import oracle as O  # noqa: E402

# ---- sections
system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
eps = -1.35
rng = np.random.default_rng(0)
pts = []
Qlo, Qhi, Pm = CD.minimum_nonnegative_Q_for_ϵ(system, eps), CD.maximum_Q_for_ϵ(system, eps), CD.maximum_P_for_ϵ(system, eps)
while len(pts) < a.n:
    Q, P = rng.uniform(Qlo, Qhi), rng.uniform(-Pm, Pm)
    if CD.minimum_ϵ_for(system, P=P, p=0, Q=Q) <= eps:
        pts.append(CD.Point(system, Q=Q, P=P, p=0, ϵ=eps))
pts = np.array(pts)
kw = dict(coef=[0, 0, 0, 1], tol=1e-10, integrator_alg="DP8", max_events=512)
CS.poincare_section_batch(system, pts[:1024], 100.0, **kw)
t0 = time.perf_counter()
r = CS.poincare_section_batch(system, pts, 1000.0, **kw)
wall = time.perf_counter() - t0
steps, ev = int(r["nsteps"].sum()), int(np.minimum(r["count"], 512).sum())
t0 = time.perf_counter()
rp = CS.poincare_section_batch(system, pts, 1000.0, packed=True, **kw)      # CSR output: no [n][max_events] padding
wall_packed = time.perf_counter() - t0
assert len(rp["t"]) == ev
m = min(256, a.n)
t0 = time.perf_counter()
et, eu, ed, cnt, ue, st, ns = O.integrate_events(O.SYS_DICKE, [1.0, 0.8, 0.8], pts[:m], 1000.0, [0, 0, 0, 1], tol=1e-10, max_events=512, fast=True)
cw = time.perf_counter() - t0
print(json.dumps({"what": "Poincare sections p=0, DOP853 tol=1e-10, t=1000, one launch", "initial_conditions": a.n, "steps": steps, "crossings": ev,
                  "kernel_ms": r["kernel_ms"], "wall_s": wall, "wall_packed_s": wall_packed, "steps_per_s_kernel": steps / (r["kernel_ms"] * 1e-3),
                  "crossings_per_s": ev / wall, "crossings_per_s_packed": ev / wall_packed,
                  "failed": int((r["status"] != 0).sum()),
                  "cpu_baseline": {"kind": "port", "cores": O.lib(True).orc_max_threads(), "initial_conditions": m, "seconds": cw,
                                   "steps_per_s": float(ns.sum() / cw), "same_counts_fraction": float((cnt == r["count"][:m]).mean())}}))

# ---- energy shell
system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
f = lambda x: [1, x[1] ** 2 + x[3] ** 2]
ESP.matrix_QP_iint_dqdp_delta(system, f=f, ϵ=-0.5, res=0.1)
t0 = time.perf_counter()
Qs, Ps, mat = ESP.matrix_QP_iint_dqdp_delta(system, f=f, ϵ=-0.5, res=a.res)
wall = time.perf_counter() - t0
info = dict(ESP.last_info)
cells = [(Q, P) for Q in Qs[::max(1, len(Qs) // 12)] for P in Ps[::max(1, len(Ps) // 12)]]
t0 = time.perf_counter()
nodes_cpu, maxdiff = 0, 0.0
for Q, P in cells:
    v = O.shell_quadrature([1.0, 1.0, 1.0], -0.5, Q, P, ["1", "q^2 + p^2"], p_res=a.res)
    if v is not None:
        i, j = int(np.argmin(np.abs(Ps - P))), int(np.argmin(np.abs(Qs - Q)))
        maxdiff = max(maxdiff, float(np.abs(mat[i, j] - np.array(v)).max()))
        A = 4 * Q * Q * (1 - (Q * Q + P * P) / 4) - (P * P + Q * Q - 2) + 2 * (-0.5)
        nodes_cpu += 2 * int(2 * np.floor((np.sqrt(A) - 1e-10) / a.res) + 1)
cw = time.perf_counter() - t0
print(json.dumps({"what": "matrix_QP∫∫dqdpδϵ, f=[1, q^2+p^2], one launch", "res": a.res, "cells": info["cells"], "cells_in_shell": info["cells_in_shell"],
                  "nodes": info["nodes"], "kernel_ms": info["kernel_ms"], "wall_s": wall, "nodes_per_s_kernel": info["nodes"] / (info["kernel_ms"] * 1e-3),
                  "shell_volume_check": float(np.nansum(mat[..., 0]) * a.res ** 2 / CD.energy_shell_volume(system, -0.5)),
                  "cpu_baseline": {"kind": "port (numpy-vectorised expression evaluation, sequential sums)", "cores": 1, "cells": len(cells),
                                   "nodes": nodes_cpu, "seconds": cw, "nodes_per_s": nodes_cpu / cw, "max_abs_diff_vs_gpu": maxdiff}}))
