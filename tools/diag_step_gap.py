#!/usr/bin/env python
"""Where does the time of a bench step go outside the ensemble kernel?  Per step: host wall clock, CUDA events around the call
on the library's stream, the library's own total (first memset to last copy) and the kernel.  Variants: with / without the
NVML clock-sampling thread of bench.py, with / without the 256 MiB L2 flush, Philox sampler / host-supplied initial conditions."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import bench  # noqa: E402

pkg = graft.load_package()
TWA, CD = pkg.TWA, pkg.ClassicalDicke
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
ctx = pkg.Context([0], stream=stream.cuda_stream)
pkg.set_default_context(ctx)
W0 = bench.WORK
system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
j = W0["j"]
ts = TWA.jrange(0, W0["t_step"], W0["t_end"])
jz = TWA.Weyl.Jz(j) / j
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 7
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
W = TWA.coherent_Wigner_HWxSU2(W0["center"], j=j, seed=W0["seed"])
e0, e1, ef = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import gc


def run(label, dist, do_flush, sampler, steps=6, freeze=False, nogc=False):
    s = None
    if freeze:
        gc.collect()
        gc.freeze()
    if nogc:
        gc.collect()
        gc.disable()
    if sampler:
        pr = torch.cuda.get_device_properties(0)
        s = bench.clocks_sampler_start("/tmp/diag_clocks.csv", 0, f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
    torch.cuda.synchronize()
    rows = []
    t_all0 = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        e0.record(stream)
        if do_flush:
            flush.zero_()
        ef.record(stream)
        t1 = time.perf_counter()
        dist.offset = 0
        TWA.average(system, observable=jz, N=N, ts=ts, distribution=dist, integrator_alg="RK4", dt=W0["dt"])
        e1.record(stream)
        torch.cuda.synchronize()
        i = TWA.last_info
        rows.append((1e3 * (time.perf_counter() - t0), e0.elapsed_time(e1), i["total_ms"], i["kernel_ms"], e0.elapsed_time(ef), 1e3 * (t1 - t0)))
    wall = 1e3 * (time.perf_counter() - t_all0) / steps
    if s:
        s.terminate()
        s.wait()
    r = np.array(rows)
    if freeze:
        gc.unfreeze()
    if nogc:
        gc.enable()
    print(f"{label:46s} wall/step {wall:8.2f} | per step (host wall, events, library total, kernel, flush on the GPU, flush call on the host) ms: "
          + "; ".join(f"{a:.1f} {b:.1f} {c:.1f} {d:.1f} {e:.2f} {f:.2f}" for a, b, c, d, e, f in r), flush=True)


for _ in range(3):
    TWA.average(system, observable=jz, N=N, ts=ts, distribution=W, integrator_alg="RK4", dt=W0["dt"])
run("philox, flush, no sampler", W, True, False)
run("philox, flush, no sampler (again), 16 steps", W, True, False, steps=16)
run("philox, flush, no sampler, gc disabled, 16 steps", W, True, False, steps=16, nogc=True)
if len(sys.argv) > 2:
    sys.exit(0)
host = torch.empty((N, 4), dtype=torch.float64).pin_memory()
hn = host.numpy()
Wh = TWA.coherent_Wigner_HWxSU2(W0["center"], j=j, seed=W0["seed"])
for s0 in range(0, N, 1 << 21):
    hn[s0:min(N, s0 + (1 << 21))] = Wh.sample_many(min(N, s0 + (1 << 21)) - s0)
D = TWA.samples_from_array(hn, j=j)
dev = torch.empty((N, 4), dtype=torch.float64, device="cuda")
for _ in range(3):
    torch.cuda.synchronize()
    e0.record(stream)
    dev.copy_(host, non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"pinned H2D of {host.numel() * 8 / 1e6:.0f} MB alone: {e0.elapsed_time(e1):.2f} ms = {host.numel() * 8 / 1e6 / e0.elapsed_time(e1):.1f} GB/s", flush=True)
del dev
run("host array (pinned), flush, no sampler", D, True, False)
os.environ["DTWA_NO_COPY_OVERLAP"] = "1"
run("host array (pinned), flush, NO upload overlap", D, True, False)
del os.environ["DTWA_NO_COPY_OVERLAP"]
run("host array (pinned), flush, overlap again", D, True, False)
