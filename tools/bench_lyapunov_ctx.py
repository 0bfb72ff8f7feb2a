#!/usr/bin/env python
"""The docs' 256 x 256 Lyapunov map through dtwa_lyapunov on a one-device and on a two-device context (the C ABI deals the
orbits round-robin over the devices of a context).  usage: python tools/bench_lyapunov_ctx.py   (needs two GPUs)"""
import sys, time, json
import numpy as np
sys.path.insert(0,'/root/repo')
import __graft_entry__ as g
pkg=g.load_package()
L=pkg._lib
CD, CS = pkg.ClassicalDicke, pkg.ClassicalSystems
system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
eps=-1.35; n=256
Qs = np.linspace(CD.minimum_nonnegative_Q_for_ϵ(system, eps), CD.maximum_Q_for_ϵ(system, eps), n)
Ps = np.linspace(0, CD.maximum_P_for_ϵ(system, eps), n)
pts = np.array([CD.Point(system, Q=Q, P=P, p=0, ϵ=eps) for Q in Qs for P in Ps if CD.minimum_ϵ_for(system, P=P, p=0, Q=Q) <= eps])
sol = CS.solver_from_kargs(dict(tol=1e-8, integrator_alg="DP8"))
one, two = L.device_context(0), L.Context([0,1])
for name,c in (("one",one),("two",two)):
    c.lyapunov(system._c_system(), sol, pts[:2048], 100.0, 5e-5, 1e-3, 100)
    t0=time.perf_counter(); r=c.lyapunov(system._c_system(), sol, pts, 100.0, 5e-5, 1e-3, 100); w=time.perf_counter()-t0
    print(json.dumps({"context": name, "points": len(pts), "kernel_ms": r[5], "wall_s": w}))
