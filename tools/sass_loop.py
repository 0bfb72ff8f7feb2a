#!/usr/bin/env python
"""Instruction mix of the hot loop of a kernel in libdicketwa.so (static SASS evidence, no GPU needed).

usage: python tools/sass_loop.py <substring of mangled kernel name> [lib]
Finds the innermost backward-branch loop that contains MUFU.RSQ64H (Dicke) or, failing that, the
loop with the most DFMA, and prints its instruction histogram.
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "dickemodel.jl_b200/csrc/libdicketwa.so"
pat = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                loops.append((tgt, addr))
    best = None
    for lo, hi in loops:
        body = [t for a, t in ins if lo <= a <= hi]
        nmufu = sum("MUFU.RSQ64H" in t for t in body)
        nd = sum(bool(re.search(r"\bD(FMA|MUL|ADD)\b", t)) for t in body)
        key = (nmufu > 0, -len(body)) if nmufu else (False, nd)
        if nd >= 20 and (best is None or (nmufu and len(body) < best[2]) or (not best[3] and nd > best[4])):
            best = (lo, hi, len(body), nmufu, nd, body)
    print("kernel:", name, " instructions:", len(ins))
    if not best:
        print("  no FP64 loop found")
        continue
    lo, hi, n, nmufu, nd, body = best
    hist = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        hist[t.split()[0].split(".")[0] + ("." + t.split()[0].split(".")[1] if t.startswith("MUFU") else "")] += 1
    fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"  hot loop {lo:#x}..{hi:#x}: {n} instructions, FP64-pipe {fp64}, MUFU.RSQ64H {nmufu}")
    print("  " + ", ".join(f"{k}:{v}" for k, v in hist.most_common()))
