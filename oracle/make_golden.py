#!/usr/bin/env python
"""Generate tests/golden/*.json -- the fixtures that pin the oracle (run in the build container).

The reference has no golden vectors and cannot run here (no Julia), so the pins are:
  truth_mpmath.json   40-digit mpmath Taylor integration (mpmath.odefun) of the reference's equations
                      of motion (src/ClassicalDicke.jl:11-17, src/ClassicalLMG.jl:8-11) from a few
                      initial points, incl. the docs' chaotic point Point(system,Q=1,P=1,p=0,eps=0.5)
                      (docs/src/TWAExamples.md:47-48).
  truth_ld.json       x87 long-double fixed-step 8th-order integration (oracle.c orc_truth_ld) of 128
                      Dicke + 64 LMG random initial conditions; validated here against the mpmath
                      rows before being written.
  scipy_pairs.json    SciPy solve_ivp DOP853 / RK45 end states + step counts for single-interval
                      runs: the oracle's adaptive controller restates SciPy's, so these must agree
                      to rounding, step counts included.
  philox_kat.json     Random123 known-answer vectors for Philox4x32-10.
Usage: python oracle/make_golden.py
"""
import json
import os
import sys

import mpmath as mp
import numpy as np
from scipy.integrate import solve_ivp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import ctypes as C  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def dicke_rhs_mp(par):
    w0, w, g = [mp.mpf(x) for x in par]

    def f(t, u):
        Q, q, P, p = u
        s = mp.sqrt(4 - P ** 2 - Q ** 2)
        return [P * w0 - g * P * q * Q / s, p * w, g * q * Q ** 2 / s - g * q * s - Q * w0, -g * Q * s - q * w]
    return f


def lmg_rhs_mp(par):
    Om, xi = [mp.mpf(x) for x in par]

    def f(t, u):
        Q, P = u
        return [P * (Om - xi * Q ** 2 / 2), Q * (xi * P ** 2 / 2 - (2 * xi + Om)) + xi * Q ** 3]
    return f


def q_of_eps(par, Q, P, p, eps):
    """src/ClassicalDicke.jl:224-228,255-273 (sgn=+), in float64 like the reference."""
    w0, w, g = par
    r = 1 - (Q ** 2 + P ** 2) / 4
    disc = 4 * g ** 2 * Q ** 2 * r - p ** 2 * w ** 2 - w * w0 * (P ** 2 + Q ** 2 - 2) + 2 * w * eps
    return (-2 * g * Q * np.sqrt(r) + np.sqrt(disc)) / w


def mp_truth(kind, par, u0, times):
    mp.mp.dps = 40
    f = dicke_rhs_mp(par) if kind == O.SYS_DICKE else lmg_rhs_mp(par)
    sol = mp.odefun(f, 0, [mp.mpf(float(x)) for x in u0], tol=mp.mpf(10) ** -32)
    rows = []
    for t in times:
        rows.append([mp.nstr(c, 25) for c in sol(t)])
    return rows


def truth_ld(kind, par, u0, ts, backwards=False, hmax=2.0 ** -9):
    nv = O.nvar(kind)
    u0 = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1, nv)
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    out = np.empty((u0.shape[0], len(ts), nv))
    par = np.ascontiguousarray(par, dtype=np.float64)
    L = O.lib()
    dp = C.POINTER(C.c_double)
    L.orc_truth_ld.argtypes = [C.c_int, dp, C.c_int, dp, C.c_long, dp, C.c_int, C.c_double, dp]
    L.orc_truth_ld(kind, O._dp(par), int(backwards), O._dp(u0), u0.shape[0], O._dp(ts), len(ts), hmax, O._dp(out))
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    dpar = [1.0, 1.0, 1.0]  # w0, w, g = 2*gc
    lpar = [1.0, -1.0]      # docs/src/ClassicalLMGExamples.md:22

    # ---- 1. mpmath truth
    docs_pt = [1.0, float(q_of_eps(dpar, 1.0, 1.0, 0.0, 0.5)), 1.0, 0.0]
    cases = [
        {"system": "dicke", "par": dpar, "u0": docs_pt, "times": [1, 2, 5, 10, 20]},
        {"system": "dicke", "par": dpar, "u0": [0.05, 0.1, -0.08, 0.02], "times": [1, 5, 10]},
        {"system": "dicke", "par": [1.0, 0.8, 0.3], "u0": [-0.7, 0.4, 0.3, -0.6], "times": [1, 5, 10]},
        {"system": "lmg", "par": lpar, "u0": [0.3, -0.2], "times": [1, 10, 50]},
    ]
    if "--skip-mpmath" in sys.argv:
        cases = []
    for c in cases:
        kind = O.SYS_DICKE if c["system"] == "dicke" else O.SYS_LMG
        print("mpmath:", c["system"], c["u0"], flush=True)
        c["u"] = mp_truth(kind, c["par"], c["u0"], c["times"])
        ld = truth_ld(kind, c["par"], c["u0"], [float(t) for t in c["times"]])[0]
        ref = np.array([[float(x) for x in row] for row in c["u"]])
        rel = np.abs(ld - ref) / np.maximum(1e-3, np.abs(ref))
        c["ld_vs_mp_maxrel"] = float(rel.max())
        print("   long-double vs mpmath max rel err:", rel.max(axis=1))
        assert rel[:3].max() < 5e-13, "long-double truth disagrees with mpmath"
    if cases:
        with open(os.path.join(GOLD, "truth_mpmath.json"), "w") as f:
            json.dump({"note": "mpmath.odefun dps=40 tol=1e-32; strings carry 25 significant digits",
                       "cases": cases}, f, indent=1)

    # ---- 2. long-double truth for random initial conditions
    rng = np.random.default_rng(12345)
    sd = O.sample(O.SAMPLER_HWXSU2, docs_pt, 1 / 100, 64, seed=7)
    sq = O.sample(O.SAMPLER_HWXSU2, [0, 0, 0, 0], 1 / 100, 32, seed=8)
    wide = np.column_stack([rng.uniform(-1.2, 1.2, 32), rng.uniform(-1, 1, 32), rng.uniform(-1.2, 1.2, 32),
                            rng.uniform(-1, 1, 32)])
    du0 = np.vstack([sd, sq, wide])
    ts = [0.0, 0.5, 1.0, 2.0, 5.0, 10.0]
    dtruth = truth_ld(O.SYS_DICKE, dpar, du0, ts, hmax=2.0 ** -10)
    dtruth2 = truth_ld(O.SYS_DICKE, dpar, du0, ts, hmax=2.0 ** -9)
    # trajectories with energy above w0 can graze the chart singularity Q^2+P^2 -> 4, where no
    # fixed-step truth is trustworthy: keep only those whose h and 2h solutions agree to 2e-13
    dis = (np.abs(dtruth - dtruth2) / np.maximum(1e-3, np.abs(dtruth))).max(axis=(1, 2))
    keep = dis < 2e-13
    print("dicke ld self-consistency: dropped", int((~keep).sum()), "of", len(keep), "max kept", dis[keep].max())
    du0, dtruth = du0[keep], dtruth[keep]
    lu0 = np.column_stack([rng.uniform(-1.3, 1.3, 64), rng.uniform(-1.3, 1.3, 64)])
    lts = [0.0, 1.0, 5.0, 20.0, 50.0]
    ltruth = truth_ld(O.SYS_LMG, lpar, lu0, lts, hmax=2.0 ** -10)
    ldis = (np.abs(ltruth - truth_ld(O.SYS_LMG, lpar, lu0, lts)) / np.maximum(1e-3, np.abs(ltruth))).max(axis=(1, 2))
    lkeep = ldis < 2e-13
    print("lmg ld self-consistency: dropped", int((~lkeep).sum()), "of", len(lkeep), "max kept", ldis[lkeep].max())
    lu0, ltruth = lu0[lkeep], ltruth[lkeep]
    with open(os.path.join(GOLD, "truth_ld.json"), "w") as f:
        json.dump({"note": "orc_truth_ld: x87 long double, fixed-step DOP853 weights, h<=2^-10, kept only where h and 2h agree to 2e-13",
                   "dicke": {"par": dpar, "ts": ts, "u0": du0.tolist(), "u": dtruth.tolist()},
                   "lmg": {"par": lpar, "ts": lts, "u0": lu0.tolist(), "u": ltruth.tolist()}}, f)

    # ---- 3. SciPy runs of the same pairs (single interval [0, T])
    def f_d(t, u):
        du, _ = O.rhs(O.SYS_DICKE, dpar, u)
        return du

    def f_l(t, u):
        du, _ = O.rhs(O.SYS_LMG, lpar, u)
        return du
    runs = []
    for name, fun, kind, par, u0s, T in (("dicke", f_d, O.SYS_DICKE, dpar, du0[[0, 1, 64, 70, 97]], 10.0),
                                         ("lmg", f_l, O.SYS_LMG, lpar, lu0[:4], 20.0)):
        for method, alg, nstage in (("DOP853", O.ALG_DOP853, 12), ("RK45", O.ALG_DP5, 6)):
            for tol in (1e-8, 1e-12):
                for u0 in u0s:
                    s = solve_ivp(fun, (0, T), u0, method=method, rtol=tol, atol=tol)
                    runs.append({"system": name, "par": par, "method": method, "tol": tol, "T": T,
                                 "u0": list(map(float, u0)), "u": list(map(float, s.y[:, -1])),
                                 "accepted": int(len(s.t) - 1), "nfev": int(s.nfev)})
    with open(os.path.join(GOLD, "scipy_pairs.json"), "w") as f:
        json.dump({"note": "scipy.integrate.solve_ivp, rtol=atol=tol, default first step", "runs": runs}, f)

    # ---- 4. Philox KAT (Random123 kat_vectors, philox4x32 10 rounds)
    kat = [{"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
           {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
           {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
            "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]}]
    for k in kat:
        assert O.philox(k["ctr"], k["key"]) == k["out"]
    with open(os.path.join(GOLD, "philox_kat.json"), "w") as f:
        json.dump(kat, f)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
