#!/usr/bin/env python
"""Generate the embedded Runge-Kutta pair coefficients used by the oracle and the CUDA library.

Source of the numbers: SciPy (BSD) `scipy/integrate/_ivp/rk.py` (RK45 = Dormand-Prince 5(4)) and
`scipy/integrate/_ivp/dop853_coefficients.py` (Hairer's DOP853 = the reference's
`integrator_alg=DP8()`).  The reference's default `TsitPap8`
(/root/reference/src/ClassicalSystems.jl:84) has no coefficient source in this image, so it is not
generated (SURVEY.md §7 hard part 8).

Tsitouras' 5(4) pair (OrdinaryDiffEq's `Tsit5()`; Ch. Tsitouras, Comput. Math. Appl. 62 (2011) 770) is not in SciPy: its
coefficients are written out below as published (16 significant digits) and CHECKED before anything is generated -- the row sums
against the nodes, all 17 order conditions up to order 5 for the weights and the 8 conditions up to order 4 for the embedded
weights must hold to 1e-14, otherwise the script stops.

Outputs
  oracle/tableaux_gen.h                      dense C arrays for the CPU restatement
  dickemodel.jl_b200/csrc/dp5_body.inc       straight-line fma code for the sm_100a kernels
  dickemodel.jl_b200/csrc/tsit5_body.inc
  dickemodel.jl_b200/csrc/dop853_body.inc

The .inc bodies expect in scope:  `double y[NV]`, `double k0[NV]` (= f(y)), `double h`, a callable
`RHS(in, out)` macro, and they define `double ynew[NV]`, `double knew[NV]` (= f(ynew)) plus the
error combination(s).  Zero coefficients are elided at generation time.
"""
import os
import numpy as np
from scipy.integrate._ivp import rk, dop853_coefficients as d8

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lit(x):
    return repr(float(x))


class ConstTable:
    """Coefficients of a body as entries of a `__constant__` array: a DFMA then takes the coefficient as a
    constant-bank operand.  Written as literals, ptxas re-materialises each one with two UMOVs per use and step
    (76 coefficients do not fit the 63 uniform registers): 150 of the 1059 instructions of a DOP853 step."""

    def __init__(self, name):
        self.name, self.vals = name, []

    def __call__(self, x):
        x = float(x)
        if x not in self.vals:
            self.vals.append(x)
        return f"{self.name.upper()}({self.vals.index(x)})"

    def definition(self):
        """the table, plus the same numbers as literals: -DDTWA_COEF_LITERALS makes every use an immediate (two UMOVs per use on
        the uniform datapath, no vector register ever holds a coefficient) -- an experiment for the register-bound adaptive kernels"""
        up = self.name.upper()
        lines = [f"__constant__ double {self.name}[{len(self.vals)}] = {{" + ", ".join(repr(v) for v in self.vals) + "};",
                 "#ifdef DTWA_COEF_LITERALS", f"#define {up}(i) {up}_##i"]
        lines += [f"#define {up}_{i} {v!r}" for i, v in enumerate(self.vals)]
        lines += ["#else", f"#define {up}(i) {self.name}[i]", "#endif"]
        return "\n".join(lines)


def tsit5_tableau():
    """A (6 x 6, strictly lower), B (6) = the seventh row (FSAL), C (6), E (7) = btilde: dt * (E . k) is the error estimate."""
    c = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
    A = np.zeros((7, 7))
    A[1, 0] = 0.161
    A[2, :2] = [-0.008480655492356989, 0.335480655492357]
    A[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
    A[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
    A[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]
    A[6, :6] = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]
    bt = np.array([-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, 0.5823571654525552,
                   -0.45808210592918697, 0.015151515151515152])
    b = A[6].copy()

    def conditions(w, order):
        Ac = A @ c
        r = [w.sum() - 1, w @ c - 1 / 2, w @ c ** 2 - 1 / 3, w @ Ac - 1 / 6, w @ c ** 3 - 1 / 4, w @ (c * Ac) - 1 / 8, w @ (A @ c ** 2) - 1 / 12,
             w @ (A @ Ac) - 1 / 24]
        if order >= 5:
            r += [w @ c ** 4 - 1 / 5, w @ (c ** 2 * Ac) - 1 / 10, w @ (c * (A @ c ** 2)) - 1 / 15, w @ (c * (A @ Ac)) - 1 / 30, w @ (Ac * Ac) - 1 / 20,
                  w @ (A @ c ** 3) - 1 / 20, w @ (A @ (c * Ac)) - 1 / 40, w @ (A @ (A @ c ** 2)) - 1 / 60, w @ (A @ (A @ Ac)) - 1 / 120]
        return np.abs(np.array(r)).max()

    assert np.abs(A.sum(1) - c).max() < 1e-14, "Tsit5: row sums"
    assert conditions(b, 5) < 1e-14, "Tsit5: order-5 conditions of the weights"
    assert conditions(b - bt, 4) < 1e-14, "Tsit5: order-4 conditions of the embedded weights"
    return A[:6, :6].copy(), b[:6].copy(), c[:6].copy(), bt


def c_array(name, arr):
    arr = np.asarray(arr, dtype=float)
    if arr.ndim == 1:
        body = ", ".join(lit(v) for v in arr)
        return f"static const double {name}[{arr.shape[0]}] = {{{body}}};\n"
    rows = ",\n  ".join("{" + ", ".join(lit(v) for v in row) + "}" for row in arr)
    return f"static const double {name}[{arr.shape[0]}][{arr.shape[1]}] = {{\n  {rows}}};\n"


def oracle_header():
    A5 = np.zeros((6, 6)); A5[:, :5] = rk.RK45.A
    out = ["/* GENERATED by tools/gen_tableaux.py from SciPy's coefficient tables - do not edit. */\n",
           "#ifndef ORACLE_TABLEAUX_GEN_H\n#define ORACLE_TABLEAUX_GEN_H\n",
           "#define DP5_STAGES 6\n#define DOP853_STAGES 12\n"]
    out.append(c_array("DP5_A", A5))
    out.append(c_array("DP5_B", rk.RK45.B))
    out.append(c_array("DP5_C", rk.RK45.C))
    out.append(c_array("DP5_E", rk.RK45.E))
    At, Bt, Ct, Et = tsit5_tableau()
    out.append(c_array("TSIT5_A", At))
    out.append(c_array("TSIT5_B", Bt))
    out.append(c_array("TSIT5_C", Ct))
    out.append(c_array("TSIT5_E", Et))
    out.append(c_array("DOP853_A", d8.A[:12, :12]))
    out.append(c_array("DOP853_B", d8.B))
    out.append(c_array("DOP853_C", d8.C[:12]))
    out.append(c_array("DOP853_E3", d8.E3))
    out.append(c_array("DOP853_E5", d8.E5))
    out.append("#endif\n")
    return "".join(out)


def lincomb(dst, coeffs, names, pre="DTWA_REAL ", lit=lit):
    """dst = sum_j coeffs[j]*names[j][i] as an fma chain (nonzero terms only)."""
    terms = [(c, n) for c, n in zip(coeffs, names) if c != 0.0]
    assert terms
    lines = []
    c0, n0 = terms[0]
    lines.append(f"{pre}{dst} = DTWA_CMUL({lit(c0)}, {n0}[i]);")
    for c, n in terms[1:]:
        lines.append(f"{dst} = fma({lit(c)}, {n}[i], {dst});")
    return lines


def body(name, A, B, E_list, n_stages):
    """A: (n_stages, n_stages) strictly lower; B: (n_stages,); E_list: [(symbol, coeffs(n_stages+1))]."""
    L = []
    lit = ConstTable(name.lower() + "_c")
    ks = [f"k{j}" for j in range(n_stages + 1)]
    nnzA = 0
    for s in range(1, n_stages):
        L.append(f"DTWA_REAL k{s}[NV];")
        L.append("{ DTWA_REAL yt[NV];")
        L.append("  _Pragma(\"unroll\") for (int i = 0; i < NV; ++i) {")
        for ln in lincomb("acc", A[s, :s], ks[:s], lit=lit):
            L.append("    " + ln)
        nnzA += int(np.count_nonzero(A[s, :s]))
        L.append("    yt[i] = fma(DTWA_STEP_H(i), acc, y[i]);")
        L.append("  }")
        L.append(f"  RHS(yt, k{s}); }}")
    L.append("DTWA_REAL ynew[NV], dy[NV];")
    L.append("_Pragma(\"unroll\") for (int i = 0; i < NV; ++i) {")
    for ln in lincomb("acc", B, ks[:n_stages], lit=lit):
        L.append("  " + ln)
    L.append("  dy[i] = acc;")
    L.append("  ynew[i] = fma(DTWA_STEP_H(i), acc, y[i]);")
    L.append("}")
    L.append(f"DTWA_REAL k{n_stages}[NV];")
    L.append(f"RHS(ynew, k{n_stages});")
    for sym, E in E_list:
        E = np.asarray(E, dtype=float)
        D = E.copy()
        D[:n_stages] -= np.asarray(B, dtype=float)   # E = B + D: reuse the weighted sum dy when D is sparser
        L.append(f"DTWA_REAL {sym}[NV];")
        L.append("_Pragma(\"unroll\") for (int i = 0; i < NV; ++i) {")
        if np.count_nonzero(D) + 1 < np.count_nonzero(E):
            L.append(f"  /* {sym} = dy + (E - B) . k : {int(np.count_nonzero(D))} terms instead of {int(np.count_nonzero(E))} */")
            L.append("  DTWA_REAL acc = dy[i];")
            for c, n in zip(D, ks):
                if c != 0.0:
                    L.append(f"  acc = fma({lit(c)}, {n}[i], acc);")
        else:
            for ln in lincomb("acc", E, ks, lit=lit):
                L.append("  " + ln)
        L.append(f"  {sym}[i] = acc;")
        L.append("}")
    L.append(f"#define {name}_KNEW k{n_stages}")
    L.append(f"/* nnz(A)={nnzA} nnz(B)={int(np.count_nonzero(B))} "
             + " ".join(f"nnz({s})={int(np.count_nonzero(E))}" for s, E in E_list) + " */")
    head = [f"/* GENERATED by tools/gen_tableaux.py ({name}) - do not edit.",
            f" * Included twice: at namespace scope with DTWA_TABLEAU_CONSTANTS defined (the coefficient table in constant",
            f" * memory), and inside a stepper (the straight-line stage code). */",
            "#ifdef DTWA_TABLEAU_CONSTANTS", lit.definition(), "#else"]
    return "\n".join(head + L + ["#endif"]) + "\n"


def main():
    with open(os.path.join(ROOT, "oracle", "tableaux_gen.h"), "w") as f:
        f.write(oracle_header())
    csrc = os.path.join(ROOT, "dickemodel.jl_b200", "csrc")
    A5 = np.zeros((6, 6)); A5[:, :5] = rk.RK45.A
    with open(os.path.join(csrc, "dp5_body.inc"), "w") as f:
        f.write(body("DP5", A5, rk.RK45.B, [("e5", rk.RK45.E)], 6))
    At, Bt, _, Et = tsit5_tableau()
    with open(os.path.join(csrc, "tsit5_body.inc"), "w") as f:
        f.write(body("TSIT5", At, Bt, [("e5", Et)], 6))
    with open(os.path.join(csrc, "dop853_body.inc"), "w") as f:
        f.write(body("DOP853", d8.A[:12, :12], d8.B, [("e5", d8.E5), ("e3", d8.E3)], 12))
    print("wrote oracle/tableaux_gen.h, csrc/dp5_body.inc, csrc/tsit5_body.inc, csrc/dop853_body.inc")


if __name__ == "__main__":
    main()
