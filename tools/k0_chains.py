#!/usr/bin/env python
"""K0c sweep: DFMA rate for (independent chains per thread) x (one-warp CTAs per SM).  Evidence for DESIGN.md section 5:
can the shape of the adaptive kernels -- 3 warps per scheduler, 4 independent chains -- keep the FP64 pipe busy at all?"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
ctx = pkg.default_context()
peak, _ = ctx.fp64_peak()
rows = []
for nreg in (1, 2, 3):
    for warps in (4, 8, 12, 16, 32):
        for chains in (1, 2, 4, 8, 16):
            tf, ms = ctx.fp64_chains(chains, warps, nreg)
            rows.append({"nreg": nreg, "warps_per_sm": warps, "chains": chains, "tflops": round(tf, 2), "frac_of_k0": round(tf / peak, 3)})
            print(rows[-1], flush=True)
print(json.dumps({"k0_tflops": peak, "rows": rows}))
