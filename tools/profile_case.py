#!/usr/bin/env python
"""A short, ncu-friendly run of the bench workload's kernel (same system, sampler, observable,
step size; fewer trajectories and a shorter time axis so that ~40 replays stay cheap).

usage: python tools/profile_case.py [--n 1048576] [--tend 5] [--alg RK4] [--reps 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--tend", type=float, default=5.0)
ap.add_argument("--alg", default="RK4")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--tstep", type=float, default=0.05, help="output-time spacing")
ap.add_argument("--block", type=int, default=0)
ap.add_argument("--ctas-per-sm", type=int, default=0)
ap.add_argument("--hist", action="store_true", help="config-3 style chain: jz-vs-t and q-vs-t histograms")
ap.add_argument("--mode", default="", choices=["", "hist2d", "animate", "variance", "survival", "chain3"])
ap.add_argument("--state", default="A", choices=["A", "B"])
ap.add_argument("--tol", type=float, default=1e-8)
ap.add_argument("--jit", type=int, default=2)
ap.add_argument("--output", default=None, help="step (default) | hermite")
ap.add_argument("--chart", default=None, help="qp (default) | bloch")
ap.add_argument("--small-bins", action="store_true", help="hist2d on an 81 x 81 grid (fits the per-CTA shared-memory tile)")
ap.add_argument("--controller", default=None, help="ode (default) | scipy")
ap.add_argument("--gains", default="", help="beta1,beta2[,qmin,qmax[,gamma]] for the OrdinaryDiffEq controller")
ap.add_argument("--w", type=float, default=1.0)
ap.add_argument("--g", type=float, default=1.0)
a = ap.parse_args()
pkg = graft.load_package()
TWA, CD = pkg.TWA, pkg.ClassicalDicke
ctx = pkg.default_context()
ctx.set_jit(a.jit)
if a.block or a.ctas_per_sm:
    ctx.set_launch(a.block, a.ctas_per_sm)
system = CD.ClassicalDickeSystem(w0=1, w=a.w, g=a.g)
j = 100
center = [0, 0, 0, 0] if a.state == "A" else [1.0, 0.31783724519578205, 1.0, 0.0]
ts = TWA.jrange(0, a.tstep, a.tend)
jz = TWA.Weyl.Jz(j) / j
ap_dt = {"RK4": 2.0 ** -7, "RK8": 0.05}
kw = dict(integrator_alg=a.alg, dt=ap_dt[a.alg]) if a.alg in ap_dt else dict(integrator_alg=a.alg, tol=a.tol)
if a.chart:
    kw["chart"] = a.chart
if a.output:
    kw["output"] = a.output
if a.alg not in ap_dt:
    if a.controller:
        kw["controller"] = a.controller
    if a.gains:
        kw.update(zip(("beta1", "beta2", "qmin", "qmax", "gamma"), (float(x) for x in a.gains.split(","))))
import hashlib
import warnings

import numpy as np


def _digest(out):   # bit-identity of the result across launch shapes
    try:
        arrs = out if isinstance(out, (list, tuple)) else [out]
        h = hashlib.sha1()
        for x in arrs:
            h.update(np.ascontiguousarray(np.asarray(x, dtype=np.float64)).tobytes())
        return h.hexdigest()[:12]
    except Exception:
        return "n/a"


warnings.simplefilter("ignore")
for rep in range(a.reps):
    W = TWA.coherent_Wigner_HWxSU2(center, j=j, seed=20261019)
    qs, ps = TWA.jrange(-4.3, 0.02, 4.3), TWA.jrange(-2.4, 0.02, 2.4)
    if a.small_bins:
        qs, ps = TWA.jrange(-4.0, 0.1, 4.0), TWA.jrange(-4.0, 0.1, 4.0)
    if a.mode == "hist2d":
        out = TWA.calculate_distribution(system, x="q", xs=qs, y="p", ys=ps, ts=ts, N=a.n, distribution=W, **kw)
    elif a.mode == "animate":
        out = TWA.calculate_distribution(system, x="q", xs=TWA.jrange(-4.3, 0.04, 4.3), y="p", ys=TWA.jrange(-2.4, 0.04, 2.4), ts=ts, N=a.n, distribution=W, animate=True, **kw)
    elif a.mode == "variance":
        out = TWA.variance(system, observable=["Q", "q", "P", "p"], ts=ts, N=a.n, distribution=W, **kw)
    elif a.mode == "survival":
        out = TWA.survival_probability(system, ts=ts, N=a.n, distribution=W, **kw)
    elif a.mode == "chain3":   # docs/src/TWAExamples.md:101-131
        mk = lambda **k: TWA.mcs_for_distributions(system, N=a.n, distribution=W, ts=ts, **k)
        out = TWA.monte_carlo_integrate(system, TWA.mcs_chain(mk(y="t", x="q", xs=qs), mk(x="t", y="p", ys=ps), mk(x="q", xs=qs, y="p", ys=ps)), **kw)
    elif a.hist:
        mk = lambda **k: TWA.mcs_for_distributions(system, N=a.n, distribution=W, ts=ts, **k)
        out = TWA.monte_carlo_integrate(system, TWA.mcs_chain(mk(x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1)), mk(y="t", x="q", xs=TWA.jrange(-4.3, 0.02, 4.3))), **kw)
    else:
        out = TWA.average(system, observable=jz, N=a.n, ts=ts, distribution=W, **kw)
    i = TWA.last_info
    steps = i["steps_accepted"] + i["steps_rejected"]
    print(f"rep {rep}: {steps} steps in {i['kernel_ms']:.3f} ms kernel = {steps / i['kernel_ms'] / 1e6:.2f} Gsteps/s; "
          f"lane efficiency {steps / max(1, i['lane_steps']):.3f}; rejected {i['steps_rejected'] / max(1, steps):.4f}; failed {i['n_failed']}; "
          f"launches {i['launches']}; jit {i['jit']}; alg {i.get('algorithm')} ctrl {i.get('controller')} {i.get('controller_gains')}; "
          f"sha1 {_digest(out)}", flush=True)
