#!/usr/bin/env python
"""Compile the run-time specialised ensemble kernel for a reducer set WITHOUT a GPU (NVRTC) and dump its
cubin + SASS (static evidence for profiles/).

usage: python tools/jit_sass.py [--alg RK4] [--sys dicke] [--expr "..."]... [--defines "-DX=1"] [--out /tmp/jit.cubin]
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--alg", default="RK4")
ap.add_argument("--sys", default="dicke")
ap.add_argument("--expr", action="append")
ap.add_argument("--defines", default="")
ap.add_argument("--out", default="/tmp/dtwa_jit.cubin")
ap.add_argument("--chart", default="qp")
ap.add_argument("--w", type=float, default=1.0)
ap.add_argument("--g", type=float, default=1.0)
a = ap.parse_args()
os.environ["DTWA_JIT_DUMP"] = a.out
if a.defines:
    os.environ["DTWA_JIT_DEFINES"] = a.defines
import __graft_entry__ as graft  # noqa: E402
pkg = graft.load_package()
lib = pkg._lib
algs = {"RK4": lib.ALG_RK4, "RK8": lib.ALG_RK8, "DP5": lib.ALG_DP5, "DP8": lib.ALG_DOP853, "DOP853": lib.ALG_DOP853}
if a.sys == "dicke":
    s = lib.System()
    s.kind = lib.SYS_DICKE
    s.params[0], s.params[1], s.params[2] = 1.0, a.w, a.g
    exprs = a.expr or ["((Q^2 + P^2)/2 - 1)*100.49875621120890/100"]
else:
    s = lib.System()
    s.kind = lib.SYS_LMG
    s.params[0], s.params[1] = 1.0, -1.0
    exprs = a.expr or ["Q", "P", "Q^2", "P^2"]
sol = lib.Solver()
sol.alg, sol.dt, sol.tol = algs[a.alg], 2.0 ** -7, 1e-8
sol.chart = lib.CHART_BLOCH if a.chart == "bloch" else lib.CHART_QP
src, nbytes = lib.jit_compile(s, sol, exprs)
print("cubin bytes:", nbytes, "->", a.out)
sass = subprocess.run(["cuobjdump", "-sass", a.out], capture_output=True, text=True).stdout
open(a.out + ".sass", "w").write(sass)
res = subprocess.run(["cuobjdump", "-res-usage", a.out], capture_output=True, text=True).stdout
print(res.strip())
