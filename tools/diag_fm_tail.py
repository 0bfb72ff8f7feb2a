#!/usr/bin/env python
"""Diagnostic: which initial conditions of tools/bench_integrate_batch.py take the most steps in dtwa_integrate_fm, and where."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
CD, CS, TWA = pkg.ClassicalDicke, pkg.ClassicalSystems, pkg.TWA
system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
W = TWA.coherent_Wigner_HWxSU2(CD.Point(system, Q=1, P=1, p=0, ϵ=0.5), j=100, seed=3)
n = 25000
U0 = np.array([W.sample() for _ in range(4096)])
U0 = np.tile(U0, (n // 4096 + 1, 1))[:n] + 1e-9 * np.random.default_rng(0).standard_normal((n, 4))
ts = np.linspace(0.0, 50.0, 51)
for name, fn in (("integrate", CS.integrate_batch), ("integrate_fm", CS.integrate_fm_batch)):
    fn(system, U0[:256], ts, tol=1e-8)
    t0 = time.perf_counter()
    r = fn(system, U0, ts, tol=1e-8)
    wall = time.perf_counter() - t0
    ns = np.asarray(r[-1])
    top = np.argsort(ns)[-5:][::-1]
    print(json.dumps({"what": name, "wall_s": wall, "mean": float(ns.mean()), "top_idx": top.tolist(), "top_steps": ns[top].tolist(),
                      "top_u0": [[float.hex(float(x)) for x in U0[i]] for i in top[:2]]}))
    if name == "integrate_fm":
        i = int(top[0])
        per_t = []
        for k in (1, 2, 5, 10, 20, 30, 40, 50):
            rr = fn(system, U0[i:i + 1], np.array([0.0, float(k)]), tol=1e-8)
            per_t.append((k, int(rr[-1][0])))
        print(json.dumps({"worst": i, "steps_to_t": per_t}))
        # the same initial condition in a full warp of copies (is it the orbit, or its neighbours in the warp?)
        rr = fn(system, np.repeat(U0[i:i + 1], 64, axis=0), ts, tol=1e-8)
        print(json.dumps({"worst_alone_x64_steps": np.asarray(rr[-1])[:4].tolist()}))
        # timing by blocks of 1024 orbits: where does the time go?
        blk = []
        for b in range(0, n, 1024):
            t0 = time.perf_counter()
            rr = fn(system, U0[b:b + 1024], ts, tol=1e-8)
            blk.append(round(1e3 * (time.perf_counter() - t0), 2))
        print(json.dumps({"ms_per_block_of_1024": blk}))
