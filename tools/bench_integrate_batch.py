#!/usr/bin/env python
"""Dense-output batches (dtwa_integrate, dtwa_integrate_fm) with an adaptive pair: n chaotic initial conditions, nt output times.
usage: python tools/bench_integrate_batch.py [--n 200000] [--nt 51]"""
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
ap.add_argument("--n", type=int, default=200000)
ap.add_argument("--nt", type=int, default=51)
a = ap.parse_args()
pkg = graft.load_package()
CD, CS, TWA = pkg.ClassicalDicke, pkg.ClassicalSystems, pkg.TWA
system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
W = TWA.coherent_Wigner_HWxSU2(CD.Point(system, Q=1, P=1, p=0, ϵ=0.5), j=100, seed=3)
U0 = np.array([W.sample() for _ in range(4096)])
U0 = np.tile(U0, (a.n // 4096 + 1, 1))[:a.n] + 1e-9 * np.random.default_rng(0).standard_normal((a.n, 4))
ts = np.linspace(0.0, 50.0, a.nt)
for name, fn, m in (("integrate", CS.integrate_batch, a.n), ("integrate_fm", CS.integrate_fm_batch, a.n // 4)):
    fn(system, U0[:1024], ts, tol=1e-8)
    t0 = time.perf_counter()
    r = fn(system, U0[:m], ts, tol=1e-8)
    wall = time.perf_counter() - t0
    steps = int(r[-1].sum())
    print(json.dumps({"what": name, "n": m, "nt": a.nt, "steps": steps, "wall_s": wall, "steps_per_s_wall": steps / wall, "ok": int((r[-2] == 0).sum())}))
