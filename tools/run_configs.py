#!/usr/bin/env python
"""Run the five BASELINE.json configs at one GPU's share of their full size and
print one JSON line each: throughput, counters, a few result values.  Evidence for profiles/.

usage: python tools/run_configs.py [--scale 1.0] [--only 1,2,3,4,5]
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
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--only", default="1,2,3,4,5")
a = ap.parse_args()
pkg = graft.load_package()
TWA, CD, CL = pkg.TWA, pkg.ClassicalDicke, pkg.ClassicalLMG
dicke = CD.ClassicalDickeSystem(w0=1, w=1, g=1)   # gamma = 2 gamma_c
j = 100
A = [0.0, 0.0, 0.0, 0.0]
B = CD.Point(dicke, Q=1, P=1, p=0, ϵ=0.5)
jz = TWA.Weyl.Jz(j) / j


def report(name, N, fn):
    fn(max(1000, N // 50))  # warm-up (and run-time specialisation) on a small ensemble
    t = time.perf_counter()
    res = fn(N)
    wall = time.perf_counter() - t
    i = TWA.last_info
    steps = i["steps_accepted"] + i["steps_rejected"]
    print(json.dumps({"config": name, "N": N, "trajectory_steps": steps, "kernel_ms": round(i["kernel_ms"], 3), "wall_s": round(wall, 3),
                      "steps_per_s_kernel": steps / (i["kernel_ms"] * 1e-3), "steps_per_s_wall": steps / wall,
                      "lane_efficiency": round(steps / max(1, i["lane_steps"]), 4), "rejected_fraction": round(i["steps_rejected"] / max(1, steps), 4),
                      "n_failed": i["n_failed"], "n_unfinished": i["n_unfinished"], "resample_rounds": i["resample_rounds"], "jit": i["jit"],
                      "result": res}), flush=True)


only = {int(x) for x in a.only.split(",")}
s = a.scale
if 1 in only:
    ts = TWA.jrange(0, 0.05, 100)
    report("1: average Jz, N=1e4, t<=100, DOP853 tol=1e-12 (reference default order/tolerance)", int(1e4),
           lambda N: [float(x) for x in TWA.average(dicke, observable=jz, N=N, ts=ts, distribution=TWA.coherent_Wigner_HWxSU2(A, j=j, seed=1))[[0, 100, 2000]]])
if 2 in only:
    ts = TWA.jrange(0, 0.05, 100)
    report("2: average Jz, N=1e7, RK4 dt=2^-7", int(1e7 * s),
           lambda N: [float(x) for x in TWA.average(dicke, observable=jz, N=N, ts=ts, distribution=TWA.coherent_Wigner_HWxSU2(A, j=j, seed=1), integrator_alg="RK4", dt=2.0 ** -7)[[0, 100, 2000]]])
if 3 in only:
    ts = TWA.jrange(0, 0.05, 40)

    def c3(N):
        W = TWA.coherent_Wigner_HWxSU2(B, j=j, seed=1)
        mk = lambda **k: TWA.mcs_for_distributions(dicke, N=N, distribution=W, ts=ts, **k)
        m1, m2 = TWA.monte_carlo_integrate(dicke, TWA.mcs_chain(mk(x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1)), mk(y="t", x="q", xs=TWA.jrange(-4.3, 0.02, 4.3))),
                                           integrator_alg="RK4", dt=2.0 ** -7)
        return {"jz_hist_mass_min": float(m1.sum(axis=0).min()), "q_hist_mass_min": float(m2.sum(axis=1).min())}
    report("3: jz-vs-t and q-vs-t histograms, one GPU's share (1.25e8) of N=1e9, t<=40, RK4 dt=2^-7", int(1.25e8 * s), c3)
if 4 in only:
    ts = TWA.jrange(0, 1, 1000)
    for alg in ("DP8", "DP5"):
        report(f"4: adaptive {alg} tol=1e-8, chaotic state B, t<=1000, one GPU's share (1e6) of N=8e6", int(1e6 * s),
               lambda N, alg=alg: [float(x) for x in TWA.average(dicke, observable=jz, N=N, ts=ts, distribution=TWA.coherent_Wigner_HWxSU2(B, j=j, seed=1), integrator_alg=alg, tol=1e-8)[[0, 10, 1000]]])
if 5 in only:
    lmg = CL.ClassicalLMGSystem(Ω=1, ξ=-1)
    ts = TWA.jrange(0, 0.1, 50)
    report("5: LMG variance [Q,P], N=1e8, t<=50, RK4 dt=2^-7", int(1e8 * s),
           lambda N: [float(x) for x in TWA.variance(lmg, observable=["Q", "P"], N=N, ts=ts, distribution=TWA.coherent_Wigner_SU2([0, 0], j=500, seed=1), integrator_alg="RK4", dt=2.0 ** -7).sum(axis=1)[[0, 100, 500]]])
