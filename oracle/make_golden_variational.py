#!/usr/bin/env python
"""Generate tests/golden/variational.json -- pins for the variational system (fundamental matrix).

The reference builds d Phi/dt = J(x) Phi from ParameterizedFunctions' SYMBOLIC Jacobian of its equations
of motion (src/ClassicalSystems.jl:121-125, src/ClassicalDicke.jl:11-17, src/ClassicalLMG.jl:8-11).  Here
the same thing is done independently of oracle.c's hand-written Jacobian: sympy differentiates the
reference's right-hand sides, and mpmath integrates the (nvar + nvar^2)-dimensional system with a
30-digit Taylor method.  Usage: python oracle/make_golden_variational.py   (about a minute)
"""
import json
import os
import sys

import mpmath as mp
import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
mp.mp.dps = 30


def build(kind, par):
    if kind == "dicke":
        Q, q, P, p = xs = sp.symbols("Q q P p")
        w0, w, g = [sp.Rational(str(v)) for v in par]
        s = sp.sqrt(4 - P ** 2 - Q ** 2)
        f = [P * w0 - g * P * q * Q / s, p * w, g * q * Q ** 2 / s - g * q * s - Q * w0, -g * Q * s - q * w]
    else:
        Q, P = xs = sp.symbols("Q P")
        Om, xi = [sp.Rational(str(v)) for v in par]
        f = [P * (Om - xi * Q ** 2 / 2), Q * (xi * P ** 2 / 2 - (2 * xi + Om)) + xi * Q ** 3]
    n = len(xs)
    J = sp.Matrix(f).jacobian(xs)
    ff = sp.lambdify(xs, f, "mpmath")
    Jf = sp.lambdify(xs, J, "mpmath")

    def rhs(t, y):
        x = y[:n]
        dx = list(ff(*x))
        Jm = Jf(*x)
        out = dx
        for c in range(n):          # Phi column-major: y[n + c*n + r]
            for r in range(n):
                out.append(sum(Jm[r, m] * y[n + c * n + m] for m in range(n)))
        return out
    return n, rhs


cases = [("dicke", [1.0, 1.0, 1.0], [1.0, 0.31783724519578224, 1.0, 0.0], [0.5, 1.0, 2.0]),
         ("dicke", [1.0, 0.8, 0.8], [0.4, -0.7, 0.3, 0.5], [0.5, 1.0, 2.0]),
         ("dicke", [1.0, 1.0, 1.0], [-1.2, 1.5, 0.2, -0.4], [0.5, 1.0, 2.0]),
         ("lmg", [1.0, -1.0], [0.5, 0.3], [1.0, 5.0]),
         ("lmg", [1.0, -1.0], [1.2, -0.6], [1.0, 5.0])]
out = []
for kind, par, u0, ts in cases:
    n, rhs = build(kind, par)
    y0 = [mp.mpf(v) for v in u0] + [mp.mpf(1 if r == c else 0) for c in range(n) for r in range(n)]
    sol = mp.odefun(rhs, 0, y0, tol=mp.mpf(10) ** -24)
    rows = []
    for t in ts:
        y = sol(t)
        rows.append({"t": t, "x": [float(v) for v in y[:n]],
                     "phi_colmajor": [float(v) for v in y[n:]]})
        print(kind, par, t, [mp.nstr(v, 12) for v in y[:n]], flush=True)
    out.append({"kind": kind, "par": par, "u0": u0, "rows": rows})
with open(os.path.join(GOLD, "variational.json"), "w") as f:
    json.dump({"note": "mpmath 30-digit Taylor integration of (x, Phi) with the sympy Jacobian of the reference's equations of motion",
               "cases": out}, f, indent=1)
print("wrote", os.path.join(GOLD, "variational.json"))
