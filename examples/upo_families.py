"""docs/src/UPOsExamples.md:27-73: the families A and B of periodic orbits from the ground state up to ϵ = 0."""
import time

import numpy as np

from _common import OUT, pkg

CD, UPOs = pkg.ClassicalDicke, pkg.UPOs
system = CD.ClassicalDickeSystem(ω=1.0, γ=1.0, **{"ω₀": 1.0})
t0 = time.perf_counter()
fam_A = UPOs.family_A(system, verbose=False)
po = fam_A(-0.5)
print(po)
print(f"  Lyapunov exponent {UPOs.lyapunov(po):.4f}, period {po.T:.4f}, action {UPOs.action(po):.4f}, <jz> {UPOs.jz_PO_average(po):.4f}")
ϵ0 = CD.minimum_energy(system)
rows, curves = [], {}
for ϵ in np.arange(ϵ0, 1e-9, 0.1):
    p = fam_A(ϵ)
    rows.append((ϵ, UPOs.energy(p), p.T))
    curves[f"A_{ϵ:.3f}"] = np.array(UPOs.QPp(p, npoints=256))
    curves[f"Amirror_{ϵ:.3f}"] = np.array(UPOs.QPp(UPOs.mirror_Qq(p), npoints=256))
pos = [fam_A(ϵ) for ϵ, _, _ in rows]
lam = UPOs.lyapunov_batch(pos)                 # all monodromy matrices in one batch per distinct period
for (ϵ, E, T), l in zip(rows, lam):
    print(f"  ϵ = {ϵ:7.3f}: energy {E:8.4f}, T = {T:.4f}, λ = {max(l, 0):.4f}")
fam_B = UPOs.family_B(system, verbose=False)
pb = fam_B(-0.5)
print("family B at -0.5:", pb, f"λ = {UPOs.lyapunov(pb):.4f}")
print(f"total {time.perf_counter() - t0:.1f} s")
np.savez(f"{OUT}/upo_families.npz", table=np.array(rows), lyapunov=lam, **curves)
