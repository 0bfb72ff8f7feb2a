"""docs/src/EnergyShellProjectionsExamples.md: classical functions integrated over the energy shell."""
import time

import numpy as np

from _common import OUT, pkg

CD, ESP = pkg.ClassicalDicke, pkg.EnergyShellProjections
system = CD.ClassicalDickeSystem(ω=1.0, γ=1.0, **{"ω₀": 1.0})
ϵ = -0.5
t0 = time.perf_counter()
Qs, Ps, mat = ESP.matrix_QP_iint_dqdp_delta(system, f=lambda x: [1, x[1] ** 2 + x[3] ** 2], ϵ=ϵ, res=0.01)
avg = ESP.energy_shell_average(system, ϵ=ϵ, f=lambda x: x[1] ** 2 + x[3] ** 2, res=0.01)
dt = time.perf_counter() - t0
vol = np.nansum(mat[..., 0]) * 0.01 ** 2
print(f"{mat.shape[0]} x {mat.shape[1]} cells in {dt:.2f} s; shell volume {vol:.4f} (closed form {CD.energy_shell_volume(system, ϵ):.4f})")
print(f"<q^2 + p^2> over the shell = {avg:.5f}")
np.savez(f"{OUT}/energy_shell.npz", Qs=Qs, Ps=Ps, mat=mat)
