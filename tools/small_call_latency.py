#!/usr/bin/env python
"""Host-to-host latency of a SMALL ensemble call (BASELINE configs[0]: 1e4 trajectories, ts=0:0.05:100, default adaptive
integrator at tol=1e-12): wall time per call against the library's own total_ms / kernel_ms, and a cProfile of the calls.
usage: python tools/small_call_latency.py [--n 10000] [--calls 30] [--jit 2]"""
import argparse
import cProfile
import io
import os
import pstats
import statistics
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--calls", type=int, default=30)
ap.add_argument("--jit", type=int, default=2)
ap.add_argument("--alg", default="TsitPap8")
ap.add_argument("--tol", type=float, default=1e-12)
a = ap.parse_args()
pkg = graft.load_package()
TWA, CD = pkg.TWA, pkg.ClassicalDicke
ctx = pkg.default_context()
ctx.set_jit(a.jit)
warnings.simplefilter("ignore")
system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
j = 100
ts = TWA.jrange(0, 0.05, 100)
jz = TWA.Weyl.Jz(j) / j
kw = dict(integrator_alg=a.alg, tol=a.tol) if a.alg not in ("RK4",) else dict(integrator_alg="RK4", dt=2.0 ** -7)


def call():
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=j, seed=7)
    t0 = time.perf_counter()
    TWA.average(system, observable=jz, N=a.n, ts=ts, distribution=W, **kw)
    return time.perf_counter() - t0, dict(TWA.last_info)


for _ in range(3):
    call()
if a.jit == 2:
    time.sleep(3.0)   # let the background specialisation finish, then measure the steady state
    call()
rows = [call() for _ in range(a.calls)]
wall = [r[0] * 1e3 for r in rows]
tot = [r[1]["total_ms"] for r in rows]
ker = [r[1]["kernel_ms"] for r in rows]
print(f"N={a.n} alg={a.alg} jit_mode={a.jit} jit_used={rows[-1][1]['jit']} launches={rows[-1][1]['launches']}: "
      f"wall median {statistics.median(wall):.2f} ms (min {min(wall):.2f}, max {max(wall):.2f}); library total_ms median {statistics.median(tot):.2f}; "
      f"kernel_ms median {statistics.median(ker):.2f}", flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(a.calls):
    call()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue()[:6000])
