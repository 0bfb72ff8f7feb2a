"""docs/src/ClassicalDickeExamples.md:82-236: Poincare surface of section p = 0 and the Lyapunov map of one energy shell."""
import time

import numpy as np

from _common import OUT, pkg

CD, CS = pkg.ClassicalDicke, pkg.ClassicalSystems
system = CD.ClassicalDickeSystem(ω=0.8, γ=0.8, **{"ω₀": 1})
ϵ = -1.35
Qs = np.linspace(CD.minimum_nonnegative_Q_for_ϵ(system, ϵ), CD.maximum_Q_for_ϵ(system, ϵ), 101)
Ps = np.linspace(0, CD.maximum_P_for_ϵ(system, ϵ), 101)
pts = np.array([CD.Point(system, Q=Q, P=P, p=0, ϵ=ϵ) for Q in Qs for P in Ps if CD.minimum_ϵ_for(system, P=P, p=0, Q=Q) <= ϵ])

t0 = time.perf_counter()
sec = CS.poincare_section_batch(system, pts[::40], 10000.0, coef=[0, 0, 0, 1], direction=1, tol=1e-10, max_events=4096, packed=True)
t_sec = time.perf_counter() - t0
print(f"Poincare surface: {len(pts[::40])} orbits to t = 10000, {len(sec['t'])} crossings of p = 0 in {t_sec:.2f} s")
g = np.abs(sec["u"][:, 3]).max()
H = CD.hamiltonian(system)
print(f"  max |p| on the section {g:.1e}, max energy drift {max(abs(H(u) - ϵ) for u in sec['u'][::997]):.1e}")

t0 = time.perf_counter()
lam, status = CS.lyapunov_spectrum_batch(system, pts, 100, tol=1e-8, λtol=5e-5, verbose=False)
t_lyap = time.perf_counter() - t0
lmax = lam.max(axis=1)
print(f"Lyapunov map: {len(pts)} orbits in {t_lyap:.2f} s; regular (λ < 0.005): {np.mean(lmax < 0.005):.2f}, max λ = {lmax.max():.3f}")
np.savez(f"{OUT}/poincare_and_lyapunov.npz", pts=pts, QP=sec["u"][:, [0, 2]], offsets=sec["offsets"], lam=lam, status=status)
