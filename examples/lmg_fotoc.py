"""docs/src/ClassicalLMGExamples.md:18-46: the FOTOC (sum of the variances of Q and P) of a coherent state centred at the
hyperbolic fixed point of the LMG model grows exponentially until it saturates."""
import time

import numpy as np

from _common import OUT, pkg

TWA, LMG = pkg.TWA, pkg.ClassicalLMG
fixed_point = LMG.Point(Q=0, P=0)
systemLMG = LMG.ClassicalLMGSystem(Ω=1, ξ=-1)
j = 500
W = TWA.coherent_Wigner_SU2(fixed_point, j=j)
ts = TWA.jrange(0, 0.1, 50)
t0 = time.perf_counter()
var = np.asarray(TWA.variance(systemLMG, observable=["Q", "P"], N=5000, ts=ts, distribution=W))
FOTOC = var.reshape(len(ts), -1).sum(axis=1)
dt = time.perf_counter() - t0
k = int(np.argmax(FOTOC > 0.1 * FOTOC.max()))
rate = np.polyfit(ts[5:k], np.log(FOTOC[5:k]), 1)[0]
print(f"N = 5000, {len(ts)} times in {dt * 1e3:.1f} ms; FOTOC(0) = {FOTOC[0]:.2e} (ħ = 1/j = {1 / j:.2e}), saturates at {FOTOC[-100:].mean():.3f}")
print(f"exponential growth rate {rate:.3f} (twice the Lyapunov exponent of the fixed point, 2 sqrt(-Ω(2ξ+Ω)) = 2 for Ω=1, ξ=-1)")
np.savez(f"{OUT}/lmg_fotoc.npz", ts=ts, FOTOC=FOTOC)
