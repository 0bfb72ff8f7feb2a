"""docs/src/TWAExamples.md:40-131: a coherent state evolved with the truncated Wigner approximation."""
import time

import numpy as np

from _common import OUT, pkg

TWA, CD = pkg.TWA, pkg.ClassicalDicke
system = CD.ClassicalDickeSystem(ω=1.0, γ=1.0, **{"ω₀": 1.0})
j = 300
x = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5)
W = TWA.coherent_Wigner_HWxSU2(x, j=j)
ts = TWA.jrange(0, 0.05, 40)
N = 200000
t0 = time.perf_counter()
jz = TWA.average(system, observable=TWA.Weyl.Jz(j) / j, distribution=W, N=N, ts=ts)
var = TWA.variance(system, observable=TWA.Weyl.Jz(j) / j, distribution=W, N=N, ts=ts)
sp = TWA.survival_probability(system, distribution=W, N=N, ts=ts)
mat = TWA.calculate_distribution(system, distribution=W, N=N, x="t", ts=ts, y=TWA.Weyl.Jz(j) / j, ys=TWA.jrange(-1.1, 0.01, 1.1))
dt = time.perf_counter() - t0
print(f"N = {N} trajectories x {len(ts)} times, 4 ensemble runs in {dt:.2f} s")
print(f"<jz>(0) = {jz[0]:.4f}, <jz>(40) = {jz[-1]:.4f}, var jz(40) = {var[-1]:.5f}, SP(0) = {sp[0]:.3f}, min SP = {sp.min():.2e}")
print("distribution of jz in time:", mat.shape, "mass per time", float(mat[:, 0].sum()))
np.savez(f"{OUT}/twa_quench.npz", ts=ts, jz=jz, var=var, sp=sp, mat=mat)
