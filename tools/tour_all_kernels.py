#!/usr/bin/env python
"""A small tour of every kernel of libdicketwa (ensemble chains with all reducer kinds under the interpreter and the
NVRTC specialisation, integrate, fundamental matrix, Lyapunov, sections, shell quadrature, sample/pdf) at sizes that
finish in seconds -- a smoke run for a new box or a new build (compute-sanitizer is closed on this pool)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
TWA, CD, CL, CS, ESP = pkg.TWA, pkg.ClassicalDicke, pkg.ClassicalLMG, pkg.ClassicalSystems, pkg.EnergyShellProjections
ctx = pkg.default_context()
dicke = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
lmg = CL.ClassicalLMGSystem(Ω=1, ξ=-1)
j = 30
B = CD.Point(dicke, Q=1, P=1, p=0, ϵ=0.5)
ts = TWA.jrange(0, 0.25, 5)
jz = TWA.Weyl.Jz(j) / j
for jit in (0, 1):
    ctx.set_jit(jit)
    W = TWA.coherent_Wigner_HWxSU2(B, j=j, seed=3)
    mk = lambda **k: TWA.mcs_for_distributions(dicke, N=3000, distribution=W, ts=ts, **k)
    chain = TWA.mcs_chain(TWA.mcs_for_averaging(dicke, observable=[jz, "q^2"], N=3000, ts=ts, distribution=W),
                          mk(x="t", y=jz, ys=TWA.jrange(-1.1, 0.1, 1.1)), mk(x="q", xs=TWA.jrange(-4, 0.5, 4), y="p", ys=TWA.jrange(-3, 0.5, 3)))
    for kw in (dict(integrator_alg="RK4", dt=2.0 ** -5), dict(integrator_alg="DP8", tol=1e-8), dict(integrator_alg="DP5", tol=1e-6)):
        out = TWA.monte_carlo_integrate(dicke, chain, **kw)
        print("chain", jit, kw["integrator_alg"], float(out[0][-1][0]), TWA.last_info["n_failed"])
    v = TWA.variance(lmg, observable=["Q", "P"], N=2000, ts=TWA.jrange(0, 0.5, 5), distribution=TWA.coherent_Wigner_SU2([0, 0], j=50, seed=1),
                     integrator_alg="RK4", dt=2.0 ** -5)
    s = TWA.survival_probability(dicke, N=1500, ts=TWA.jrange(0, 0.5, 3), distribution=TWA.coherent_Wigner_HWxSU2(B, j=j, seed=4), tol=1e-8)
    print("variance/survival", jit, float(v[-1][0]), float(s[-1]))
ctx.set_jit(2)
u0 = np.array([B, [0.4, -0.7, 0.3, 0.5], [-1.2, 1.5, 0.2, -0.4]])
print("integrate", CS.integrate_batch(dicke, u0, [0.0, 1.0, 2.0], tol=1e-10)[0][:, -1, 0])
print("fm", CS.integrate_fm_batch(dicke, u0, [0.0, 1.0], tol=1e-10)[1][0, -1, 0])
print("lyap", CS.lyapunov_spectrum_batch(dicke, u0, 10.0, tol=1e-8, maxiters=20)[0][:, 0])
print("lyap lmg", CS.lyapunov_spectrum_batch(lmg, [[0.5, 0.3], [1.0, 0.2], [0.1, 0.1]], 10.0, integrator_alg="RK4", dt=0.01, maxiters=5)[0][:, 0])
print("events", CS.poincare_section_batch(dicke, u0, 20.0, coef=[0, 0, 0, 1], tol=1e-9, max_events=16)["count"])
print("shell", ESP.energy_shell_average(dicke, ϵ=-0.5, f=lambda x: x[1] ** 2, res=0.25))
print("sample/pdf", ctx.pdf(TWA.coherent_Wigner_HWxSU2(B, j=j, seed=1)._c_sampler(0), ctx.sample(TWA.coherent_Wigner_HWxSU2(B, j=j, seed=1)._c_sampler(0), 100)).mean())
