#!/usr/bin/env python
"""K0d sweep: K0's DFMA chains with other instructions woven in -- does an instruction of another pipe take a DFMA issue slot?
Evidence for DESIGN.md section 5 (what bounds the adaptive kernels at ~70 % FP64-pipe activity)."""
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
kinds = ((0, "iadd"), (1, "ffma"), (2, "mufu"), (3, "imad"), (4, "lop3"))
if len(sys.argv) > 1:
    kinds = tuple(k for k in kinds if k[1] in sys.argv[1:])
for kind, name in kinds:
    for warps in (4, 12, 32):
        for alu in (0, 2, 4, 8, 16):
            tf, ms = ctx.fp64_mix(alu, kind, warps)
            rows.append({"other": name, "other_per_8_dfma": alu, "warps_per_sm": warps, "dfma_tflops": round(tf, 2), "frac_of_k0": round(tf / peak, 3)})
            print(rows[-1], flush=True)
print(json.dumps({"k0_tflops": peak, "rows": rows}))
