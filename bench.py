#!/usr/bin/env python
"""bench.py -- the headline measurement: FP64 trajectory-steps/s of the TWA ensemble on B200.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun launches N ranks)
    python bench.py --impl reference --gpus N --steps K ...  (the restated reference CPU path)

Workload (BASELINE.json configs[1]): Dicke w = w0 = 1, gamma = 2 gamma_c = 1,
j = 100, coherent state of the normal-phase ground state x0 = [0,0,0,0] (the quench), 10^7
trajectories PER GPU, ts = 0:0.05:100 (2001 output times), fixed-step RK4 with dt <= 2^-7: the library
steps exactly onto every output time, so an output interval of 0.05 takes 7 equal sub-steps of 0.05/7
(14000 RK4 steps per trajectory; SURVEY 8d-2 counts 12800 for a step of exactly 2^-7), observable
Weyl.Jz(j)/j averaged at every output time.  One bench "step" = one whole ensemble.

  value     device-resident leg: initial conditions are drawn inside the kernel (Philox), timed with
            CUDA events on the stream the kernels run on, all K steps in one bracket.
  e2e       the same ensemble through the public API (TWA.average) with HOST-supplied initial
            conditions in pinned memory: H2D of N x 32 B and D2H of the per-time sums inside the
            timed region.
  roofline  FP64 pipe: achieved = 136 algorithmic flops/step x steps / kernel time, peak = the DFMA
            micro-benchmark (K0) run in this process (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline  oracle/oracle.c (the restated reference CPU path; Julia is not in the image) on all
            host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_STEP = {"dicke_rk4": 136.0}      # SURVEY.md section 8d / BASELINE.md section 4
FP64_INSTR_PER_STEP = {"dicke_rk4": 88.0}  # SASS of the run-time specialised hot loop for w = g = 1: 64 DFMA + 24 DMUL (96 for general w and g); tools/jit_sass.py dumps it

WORK = {"system": "dicke", "par": [1.0, 1.0, 1.0], "j": 100, "center": [0.0, 0.0, 0.0, 0.0], "t_end": 100, "t_step": 0.05,
        "dt": 2.0 ** -7, "seed": 20261019}


def workload_name(n_per_gpu):
    return (f"configs[1]: Dicke TWA quench w=w0=1 g=2gc j=100, coherent state x0=0, {n_per_gpu:.0e} trajectories/GPU, "
            f"ts=0:0.05:100 (nt=2001), RK4 with 7 sub-steps of 0.05/7 per output interval (dt <= 2^-7; 14000 steps/trajectory), <Jz>/j")


class _NvmlSampler:
    """SM clock, power and clock-event reasons of one GPU every 200 ms from a thread of this process (NVML through pynvml:
    ~50 us per sample).  Writes the rows `nvidia-smi --query-gpu=... -lms 200` would write, for clocks_summary().  A separate
    nvidia-smi process polling during the timed region cost the device leg 1-17 ms per 0.8 s step depending on the box
    (profiles/r2_bench_n1_*.json: ms_per_step - kernel_ms_per_step), which the leg without a sampler never showed."""

    def __init__(self, path, device_index, pci_bus_id):
        import threading
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode()) if pci_bus_id else pynvml.nvmlDeviceGetHandleByIndex(device_index)
        self.idx, self.f, self.stop = device_index, open(path, "w"), threading.Event()
        self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        names = (nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown, nv.nvmlClocksEventReasonSwThermalSlowdown,
                 nv.nvmlClocksEventReasonSwPowerCap)
        while not self.stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                flags = ",".join("Active" if (r & m) else "Not Active" for m in names)
                self.f.write(f"{self.idx},{sm},{self.smax},{pw:.2f},0x{r:016x},{flags}\n")
                self.f.flush()
            except Exception:
                pass
            self.stop.wait(0.2)

    def terminate(self):
        self.stop.set()

    def wait(self):
        self.t.join(timeout=2.0)
        self.f.close()


def clocks_sampler_start(path, device_index=0, pci_bus_id=None):
    try:
        return _NvmlSampler(path, device_index, pci_bus_id)
    except Exception:
        pass
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except OSError:
        return None


def clocks_summary(path, device_index):
    sm, smax, reasons = [], 0.0, set()
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or f[0] != str(device_index):
                continue
            sm.append(float(f[1]))
            smax = max(smax, float(f[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except (OSError, ValueError):
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    busy = [x for x in sm if x >= 0.5 * max(sm)]
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference(N, threads, ts):
    """the restated reference CPU path on `N` trajectories of the bench workload"""
    from oracle import oracle as O
    sp = O.make_sampler(O.SAMPLER_HWXSU2, WORK["center"], 1.0 / WORK["j"], seed=WORK["seed"])
    c = float(np.sqrt(WORK["j"] * (WORK["j"] + 1)))
    t0 = time.perf_counter()
    r = O.ensemble(O.SYS_DICKE, WORK["par"], sp, N, ts, values=[(2, 0, c)], alg=O.ALG_RK4, dt=WORK["dt"], workers=threads, fast=True)
    dt = time.perf_counter() - t0
    return r, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from dickemodel_jl_b200_loader import TWA  # noqa: F401  (only for jrange)
    ts = np.asarray(TWA.jrange(0, WORK["t_step"], WORK["t_end"]))
    threads = host_threads()
    pilot_n = 64 * threads
    cpu_reference(threads, threads, ts)   # load the library, spin up the OpenMP team
    r, dt = cpu_reference(pilot_n, threads, ts)
    rate = r["steps_accepted"] / dt
    n_step = max(threads, int(rate * args.ref_seconds / (14000.0)))
    for _ in range(args.warmup):
        cpu_reference(max(threads, n_step // 8), threads, ts)
    tot_steps, tot_t = 0, 0.0
    for _ in range(args.steps):
        r, dt = cpu_reference(n_step, threads, ts)
        tot_steps += r["steps_accepted"]
        tot_t += dt
    value = tot_steps / tot_t
    line = {"impl": "reference", "metric": "FP64 trajectory-steps/s", "value": value, "unit": "trajectory-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n_traj), "note": "CPU arm runs a bounded sample of the same ensemble per step"},
            "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{n_step} trajectories x 14000 RK4 steps per step, oracle/oracle.c -O3 -march=x86-64-v3, OpenMP {threads} threads "
                                       "(restated reference CPU path; Julia unavailable)"},
            "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there at the C level (NCCL prints its version line on the first
    communicator), so fd 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-traj", type=float, default=1e7, help="trajectories per GPU per step")
    ap.add_argument("--ref-seconds", type=float, default=6.0, help="CPU seconds per step of the reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU seconds for the cpu_baseline sample")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras-scale", type=float, default=1.0, help="scale the ensemble sizes of the extras (quick checks)")
    args = ap.parse_args()
    args.n_traj = int(args.n_traj)
    claim_stdout()

    import __graft_entry__ as graft
    pkg = graft.load_package()
    sys.modules["dickemodel_jl_b200_loader"] = pkg
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side waits (a rank parked in an NCCL barrier keeps its GPU spinning)
    TWA, CD = pkg.TWA, pkg.ClassicalDicke
    stream = torch.cuda.current_stream()
    ctx = pkg.Context([local], stream=stream.cuda_stream)
    pkg.set_default_context(ctx)
    if args.block or args.ctas_per_sm:
        ctx.set_launch(args.block, args.ctas_per_sm)
    allreduce = pkg.distributed.torch_allreduce(local) if world > 1 else None

    def distributed_mode(on=True):
        pkg.set_distributed(rank if on else 0, world if on else 1, allreduce if on else None)
    distributed_mode(True)

    system = CD.ClassicalDickeSystem(w0=WORK["par"][0], w=WORK["par"][1], g=WORK["par"][2])
    j = WORK["j"]
    ts = TWA.jrange(0, WORK["t_step"], WORK["t_end"])
    jz = TWA.Weyl.Jz(j) / j
    N_total = args.n_traj * world
    solver = dict(integrator_alg="RK4", dt=WORK["dt"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def allmax(vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def gather(val):
        t = torch.tensor([float(val)], dtype=torch.float64, device="cuda")
        if world == 1:
            return [float(val)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(x[0]) for x in out]

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        """fn() bracketed by barrier + synchronize on both sides, CUDA events on the library's stream; max over ranks (ms)"""
        barrier()
        e0.record(stream)
        out = fn()
        e1.record(stream)
        barrier()
        return out, allmax([e0.elapsed_time(e1)])[0]

    # ---- roofline denominator: K0 DFMA chains ----
    # (two launch shapes of the same DFMA chains -- 8 CTAs of 256 threads per SM, and 32 one-warp CTAs per SM with longer
    #  loops; the second reads ~7 % higher on B200 -- the denominator is the BEST of the two)
    k0a_tflops, k0a_ms = ctx.fp64_peak()
    k0c_tflops, k0c_ms = ctx.fp64_chains(8, 32, 1)
    peak_tflops, k0_ms = max((k0a_tflops, k0a_ms), (k0c_tflops, k0c_ms))

    # ---- device-resident leg ----
    W = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])

    def step_device():
        flush.zero_()
        W.offset = 0   # every step integrates the same ensemble (trajectory indices 0..N-1)
        out = TWA.average(system, observable=jz, N=N_total, ts=ts, distribution=W, **solver)
        return out, dict(TWA.last_info)

    for _ in range(args.warmup):
        step_device()
    barrier()
    clk_path = os.path.join(ROOT, "gpurun_out", f"clocks_rank{rank}.csv")
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    try:
        pr = torch.cuda.get_device_properties(local)
        pci = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    except Exception:
        pci = None
    sampler = clocks_sampler_start(clk_path, local, pci) if rank == 0 else None
    # Python's cyclic collector is switched off inside the timed loops, as `timeit` does (a collection between two calls of a
    # synchronous API leaves the GPU idle).  It does not remove the sporadic late calls of a busy host (profiles/r2_step_gap_call32.log:
    # single steps 10-60 ms longer with an unchanged library time, with and without the collector): those stay in the number.
    gc.collect()
    gc.disable()
    e0.record(stream)
    steps_done, kernel_ms, launches, last = 0, 0.0, 0, None
    step_wall_ms, step_lib_ms = [], []
    for _ in range(args.steps):
        t_step = time.perf_counter()
        last, info = step_device()
        step_wall_ms.append(round(1e3 * (time.perf_counter() - t_step), 2))   # (host clock, this rank: diagnostic only)
        step_lib_ms.append(round(info.get("total_ms", 0.0), 2))
        steps_done += info["steps_accepted"]       # all-reduced: whole job
        kernel_ms += info["kernel_ms"]
        launches += info["launches"]
    e1.record(stream)
    gc.enable()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    if sampler:
        sampler.terminate()
        sampler.wait()
    elapsed_ms, kernel_ms = allmax([elapsed_ms, kernel_ms])
    value = steps_done / (elapsed_ms * 1e-3)
    per_gpu_kernel_steps = steps_done / world                       # steps one GPU's kernels processed
    kern_rate = per_gpu_kernel_steps / (kernel_ms * 1e-3)           # per-GPU steps/s inside the ensemble kernel
    achieved_tflops = kern_rate * FLOPS_PER_STEP["dicke_rk4"] / 1e12
    clocks = clocks_summary(clk_path, local) if rank == 0 else {}

    # ---- end-to-end leg: public API, host-supplied initial conditions from pinned memory ----
    e2e = None
    if not args.no_e2e:
        n_loc = args.n_traj
        host = torch.empty((N_total, 4), dtype=torch.float64).pin_memory()
        # (not timed) fill this rank's block with the same Philox samples the device leg drew
        lo, hi = pkg.distributed.shard_bounds(N_total, rank, world)
        Wh = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
        Wh.offset = lo
        host_np = host.numpy()
        chunk = 1 << 21
        for s in range(lo, hi, chunk):
            host_np[s:min(hi, s + chunk)] = Wh.sample_many(min(hi, s + chunk) - s)
        D = TWA.samples_from_array(host_np, j=j)

        def step_e2e():
            flush.zero_()
            D.offset = 0
            out = TWA.average(system, observable=jz, N=N_total, ts=ts, distribution=D, **solver)
            return out, dict(TWA.last_info)

        for _ in range(max(1, min(args.warmup, 3))):
            step_e2e()
        barrier()
        gc.collect()
        gc.disable()
        e0.record(stream)
        st, e2e_out = 0, None
        n_e2e = max(1, args.steps)
        for _ in range(n_e2e):
            e2e_out, info = step_e2e()
            st += info["steps_accepted"]
        e1.record(stream)
        gc.enable()
        barrier()
        e2e_ms = allmax([e0.elapsed_time(e1)])[0]
        e2e = {"value": st / (e2e_ms * 1e-3), "unit": "trajectory-steps/s",
               "h2d_bytes_per_step": int(n_loc * 4 * 8 * world + len(ts) * 8 * world), "d2h_bytes_per_step": int(len(ts) * 8 * world + len(ts) * 8 * world),
               "steps": n_e2e, "ms_per_step": e2e_ms / n_e2e,
               "note": "TWA.average with distribution=samples_from_array(pinned host u0[N][4]); H2D of the initial conditions and D2H of the per-time sums inside the timed region"}
        # same samples, same kernel: the two legs must agree to rounding
        e2e["max_abs_diff_vs_device_leg"] = float(np.max(np.abs(e2e_out - last))) if world == 1 else None
        del host, D, host_np

    # ---- CPU baseline (rank 0, N = 1 only): the restated reference CPU path on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        tsn = np.asarray(ts)
        cpu_reference(threads, threads, tsn)
        r, dtp = cpu_reference(32 * threads, threads, tsn)
        rate = r["steps_accepted"] / dtp
        n_cpu = max(threads, int(rate * args.cpu_seconds / 14000.0))
        r, dtc = cpu_reference(n_cpu, threads, tsn)
        r1, dt1 = cpu_reference(max(1, n_cpu // (threads * 4)), 1, tsn)
        cpu = {"value": r["steps_accepted"] / dtc, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
               "sample": f"{n_cpu} trajectories of the same workload (x 14000 RK4 steps), oracle/oracle.c -O3 -march=x86-64-v3 with OpenMP over "
                         f"{threads} threads: restated reference CPU path (Julia is not in the image); single-thread: "
                         f"{r1['steps_accepted'] / dt1:.3e} steps/s",
               "single_thread_value": r1["steps_accepted"] / dt1, "seconds": dtc}
        # Monte-Carlo agreement of the two ensembles (same Philox samples for the first n_cpu trajectories)
        cpu["max_abs_diff_mean_jz_vs_gpu"] = float(np.max(np.abs(r["sums"][:, 0] / j / n_cpu - last)))
        cpu["mc_standard_error_bound"] = float(4.0 / np.sqrt(n_cpu))

    # DRAM bytes of one launch of the dominant kernel, from the committed `ncu --set full` capture of this
    # same workload (profiles/traffic_rk4_bench.json, written by tools/ncu_summary.py --traffic-json)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_rk4_bench.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("n_traj", 0)) == args.n_traj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_src = tj.get("source")

    # ==== extras: the other BASELINE.json configs, at EVERY world size (every rank takes part) ====
    extras = {}
    scale = args.extras_scale
    Bpt = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5)      # the docs' chaotic point (state B of SURVEY 8d)
    hist_check = {}

    def hist_chain(N, W_):
        """SURVEY 8d-3 / docs/src/TWAExamples.md:75-83,101-117: Jz/j-vs-t and q-vs-t histograms over the same trajectories"""
        tsh = TWA.jrange(0, 0.05, 40)
        mk = lambda **k: TWA.mcs_for_distributions(system, N=N, distribution=W_, ts=tsh, **k)   # noqa: E731
        m = TWA.monte_carlo_integrate(system, TWA.mcs_chain(mk(x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1)), mk(y="t", x="q", xs=TWA.jrange(-4.3, 0.02, 4.3))),
                                      integrator_alg="RK4", dt=WORK["dt"])
        return [np.rint(np.asarray(x) * N).astype(np.uint64) for x in m], dict(TWA.last_info)   # back to the integer counts

    def leg(name, fn):
        """never lose the headline over an extra; an exception is recorded (it would be raised on every rank alike)"""
        try:
            fn()
        except Exception as e:   # noqa: BLE001
            extras[name] = {"error": repr(e)}
        finally:
            ctx.set_jit(2)

    def warm(fn):
        """warm-up call with the run-time specialisation forced, so that the NVRTC compile (about a second per new kernel)
        happens here and not inside the timed call that follows"""
        ctx.set_jit(1)
        try:
            return fn()
        finally:
            ctx.set_jit(2)

    def leg_hist():
        n_per = max(1024, int(1.25e8 * scale))
        Nh = n_per * world
        warm(lambda: hist_chain(max(1024, Nh // 64), TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=WORK["seed"])))   # warm-up (and run-time specialisation)
        (counts, ih), wall_ms = timed(lambda: hist_chain(Nh, TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=WORK["seed"])))
        kms = gather(ih["kernel_ms"])
        # integer bit-exactness across the GPU count: the same index range [0, Nc), sharded over all ranks + all-reduce,
        # against ONE rank integrating all of it (world = 1: against two half-ranges run one after the other and added)
        Nc = max(2048, int((1 << 22) * min(1.0, scale)))
        dist_counts, _ = hist_chain(Nc, TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=WORK["seed"]))
        hist_check["Nc"], hist_check["counts"] = Nc, dist_counts
        equal = None
        if rank == 0:
            distributed_mode(False)
            try:
                if world > 1:
                    one, _ = hist_chain(Nc, TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=WORK["seed"]))
                else:
                    Wa = TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=WORK["seed"])
                    h1, _ = hist_chain(Nc // 2, Wa)                    # indices [0, Nc/2)
                    h2, _ = hist_chain(Nc - Nc // 2, Wa)               # indices [Nc/2, Nc): the distribution's offset advanced
                    one = [x + y for x, y in zip(h1, h2)]
                equal = bool(all(np.array_equal(x, y) for x, y in zip(one, dist_counts)))
            finally:
                distributed_mode(True)
        host_barrier()
        att = ih["steps_accepted"] + ih["steps_rejected"]
        extras["config2_histograms"] = {
            "workload": f"configs[2]: mcs_chain of Jz/j-vs-t (221 bins) and q-vs-t (431 bins) histograms, state B, ts=0:0.05:40 (nt=801), RK4 dt<=2^-7, "
                        f"{n_per:.3g} trajectories/GPU (weak scaling; 1e9 at 8 GPUs)",
            "trajectories_total": Nh, "steps_per_s": att / (max(kms) * 1e-3), "steps_per_s_wall": att / (wall_ms * 1e-3),
            "kernel_ms_max": max(kms), "kernel_ms_min": min(kms), "wall_ms": wall_ms, "n_failed": ih["n_failed"], "resample_rounds": ih["resample_rounds"],
            "histogram_mass": [float(c.sum()) / (Nh * 801) for c in counts],
            "allreduce_bytes": int(8 * (801 * 221 + 801 * 431 + 801 + 8)),
            "counts_check": {"trajectories": Nc, "bit_equal": equal,
                             "against": "one rank integrating the whole index range" if world > 1 else "two half index ranges run one after the other and added"}}

    def leg_adaptive():
        tsa = TWA.jrange(0, 1, 1000)
        Na = max(world * 32, int(8e6 * scale))      # FIXED total: strong scaling
        for alg, label in (("TsitPap8", "adaptive_default_pi"), ("DP8", "adaptive_dp8"), ("DP5", "adaptive_dp5")):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                warm(lambda: TWA.average(system, observable=jz, N=max(world * 32, Na // 64), ts=tsa, distribution=TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=1),
                                         integrator_alg=alg, tol=1e-8))     # warm-up (and run-time specialisation)
                _, wall_ms = timed(lambda: TWA.average(system, observable=jz, N=Na, ts=tsa, distribution=TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=1),
                                                       integrator_alg=alg, tol=1e-8))
            ia = dict(TWA.last_info)
            kms = gather(ia["kernel_ms"])
            att = ia["steps_accepted"] + ia["steps_rejected"]
            one_gpu_rate = None
            extras[label] = {
                "workload": f"configs[3]: chaotic state B, tol=1e-8, ts=0:1:1000, N={Na:.3g} trajectories in total (strong scaling), integrator_alg={alg}",
                "algorithm_run": ia.get("algorithm"), "controller": ia.get("controller"), "controller_gains": ia.get("controller_gains"),
                "attempted_steps": att, "attempted_steps_per_s": att / (max(kms) * 1e-3), "attempted_steps_per_s_wall": att / (wall_ms * 1e-3),
                "accepted_steps_per_s": ia["steps_accepted"] / (max(kms) * 1e-3),
                "lane_efficiency": att / max(1, ia["lane_steps"]), "rejected_fraction": ia["steps_rejected"] / max(1, att),
                "kernel_ms_per_rank": kms, "kernel_ms_max": max(kms), "kernel_ms_min": min(kms), "limiting_rank": int(np.argmax(kms)),
                "imbalance_max_over_mean": max(kms) / (sum(kms) / len(kms)), "wall_ms": wall_ms,
                "flops_per_attempt": {"DOP853": 876, "DP5": 366}.get(ia.get("algorithm")), "specialised": ia["jit"],
                "roofline_frac_of_k0": (att / world / (max(kms) * 1e-3)) * {"DOP853": 876, "DP5": 366}.get(ia.get("algorithm"), 0) / (peak_tflops * 1e12),
                "n_failed": ia["n_failed"], "n_unfinished": ia["n_unfinished"]}
            del one_gpu_rate

    def leg_general():
        """the headline workload away from the unit-parameter specialisation (w = g = 1 saves 8 of 96 FP64 instructions/step)"""
        sysg = CD.ClassicalDickeSystem(w0=1.0, w=0.8, g=0.66)
        Wg = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
        warm(lambda: TWA.average(sysg, observable=jz, N=max(world * 1024, N_total // 64), ts=ts, distribution=Wg, **solver))
        Wg.offset = 0
        _, wall_ms = timed(lambda: TWA.average(sysg, observable=jz, N=N_total, ts=ts, distribution=Wg, **solver))
        ig = dict(TWA.last_info)
        km = max(gather(ig["kernel_ms"]))
        extras["general_parameters"] = {"workload": "the headline workload at w0=1, w=0.8, g=0.66: no unit-parameter specialisation (96 FP64 instr/step)",
                                        "steps_per_s": ig["steps_accepted"] / (km * 1e-3), "steps_per_s_wall": ig["steps_accepted"] / (wall_ms * 1e-3),
                                        "kernel_ms": km, "specialised": ig["jit"]}

    def leg_bloch():
        """opt-in chart="bloch" (the same equations of motion in (jx,jy,jz,q,p): bilinear, no square root) on the headline
        workload and on configs[3] -- NOT the parity path: it leaves the reference's step sequence"""
        Wb = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
        warm(lambda: TWA.average(system, observable=jz, N=max(world * 1024, N_total // 64), ts=ts, distribution=Wb, chart="bloch", **solver))
        Wb.offset = 0
        outb, wall_ms = timed(lambda: TWA.average(system, observable=jz, N=N_total, ts=ts, distribution=Wb, chart="bloch", **solver))
        ib = dict(TWA.last_info)
        km = max(gather(ib["kernel_ms"]))
        res = {"workload": "the headline workload with chart='bloch' (59 FP64 instr/step instead of 88)",
               "rk4_steps_per_s": ib["steps_accepted"] / (km * 1e-3), "rk4_steps_per_s_wall": ib["steps_accepted"] / (wall_ms * 1e-3), "rk4_kernel_ms": km,
               "max_abs_diff_mean_jz_vs_qp_chart": float(np.max(np.abs(outb - last))),
               "max_abs_diff_mean_jz_vs_qp_chart_t_le_2": float(np.max(np.abs(outb - last)[:41])),
               "mc_standard_error": float(0.3 / np.sqrt(N_total)),
               "note": "same Philox samples as the headline leg; the two charts differ by the RK4 truncation error (t <= 2: see above), which the chaotic "
                       "dynamics amplifies until, at late times, the two ensembles are independent at the level of the Monte-Carlo standard error"}
        tsa = TWA.jrange(0, 1, 1000)
        Na = max(world * 32, int(8e6 * scale))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            warm(lambda: TWA.average(system, observable=jz, N=max(world * 32, Na // 64), ts=tsa, distribution=TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=1),
                                     tol=1e-8, chart="bloch"))
            _, wall_a = timed(lambda: TWA.average(system, observable=jz, N=Na, ts=tsa, distribution=TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=1), tol=1e-8, chart="bloch"))
        ia = dict(TWA.last_info)
        kms = gather(ia["kernel_ms"])
        att = ia["steps_accepted"] + ia["steps_rejected"]
        res["configs3_default_pi"] = {"attempted_steps": att, "attempted_steps_per_s": att / (max(kms) * 1e-3), "kernel_ms_max": max(kms), "wall_ms": wall_a,
                                      "rejected_fraction": ia["steps_rejected"] / max(1, att), "lane_efficiency": att / max(1, ia["lane_steps"]),
                                      "n_failed": ia["n_failed"],
                                      "time_vs_qp_chart": (max(kms) / extras["adaptive_default_pi"]["kernel_ms_max"]) if "kernel_ms_max" in extras.get("adaptive_default_pi", {}) else None}
        extras["bloch_chart"] = res

    def leg_inproc():
        """the in-process N-device context (dtwa_create(ndev = world) + ncclCommInitAll + grouped ncclAllReduce: what the Julia
        shim uses), driven by rank 0 alone while the other ranks wait on the host"""
        if world == 1:
            return
        res = None
        if rank == 0:
            distributed_mode(False)
            try:
                ctxN = pkg.Context(list(range(world)))
                pkg.set_default_context(ctxN)
                Wn = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
                TWA.average(system, observable=jz, N=max(world * 1024, N_total // 64), ts=ts, distribution=Wn, **solver)
                Wn.offset = 0
                t0 = time.perf_counter()
                outN = TWA.average(system, observable=jz, N=N_total, ts=ts, distribution=Wn, **solver)
                wall = time.perf_counter() - t0
                iN = dict(TWA.last_info)
                cN, _ = hist_chain(hist_check["Nc"], TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=WORK["seed"]))
                res = {"nranks": world, "steps_per_s": iN["steps_accepted"] / (iN["kernel_ms"] * 1e-3), "steps_per_s_wall": iN["steps_accepted"] / wall,
                       "kernel_ms": iN["kernel_ms"], "max_abs_diff_vs_torchrun_leg": float(np.max(np.abs(outN - last))),
                       "hist_counts_bit_equal_to_torchrun": bool(all(np.array_equal(x, y) for x, y in zip(cN, hist_check["counts"]))),
                       "note": "dtwa_create(ndev) context in ONE process (rank 0): contiguous index blocks, one grouped ncclAllReduce of the arenas"}
                ctxN.close()
            finally:
                pkg.set_default_context(ctx)
                distributed_mode(True)
        host_barrier()
        if res:
            extras["inproc_ctx"] = res

    if not args.no_extras:
        leg("config2_histograms", leg_hist)
        leg("adaptive", leg_adaptive)
        leg("general_parameters", leg_general)
        leg("bloch_chart", leg_bloch)
        leg("inproc_ctx", leg_inproc)

    # ---- one GPU only, for the record: configs[0] next to the CPU port, and small instances of SURVEY 8f ranks 2-4 ----
    if rank == 0 and world == 1 and not args.no_cpu and not args.no_extras:
        try:
            # BASELINE.json configs[0] (the CPU-runnable case): N = 1e4, adaptive 8th order at the reference's default
            # tol = 1e-12, GPU next to the oracle port on all host cores, same Philox samples
            from oracle import oracle as O
            import warnings
            ts0 = TWA.jrange(0, WORK["t_step"], WORK["t_end"])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for _ in range(2):
                    W0 = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
                    t0 = time.perf_counter()
                    g0 = TWA.average(system, observable=jz, N=10000, ts=ts0, distribution=W0)
                    wall0 = time.perf_counter() - t0
            i0 = dict(TWA.last_info)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for _ in range(2):
                    W0 = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
                    t0 = time.perf_counter()
                    g0h = TWA.average(system, observable=jz, N=10000, ts=ts0, distribution=W0, output="hermite")
                    wall0h = time.perf_counter() - t0
            i0h = dict(TWA.last_info)
            sp0 = O.make_sampler(O.SAMPLER_HWXSU2, WORK["center"], 1.0 / j, seed=WORK["seed"])
            cj = float(np.sqrt(j * (j + 1)))
            thr = host_threads()
            t0 = time.perf_counter()
            r0 = O.ensemble(O.SYS_DICKE, WORK["par"], sp0, 10000, np.asarray(ts0), values=[(2, 0, cj)], alg=O.ALG_DOP853, tol=1e-12,
                            workers=thr, fast=True, controller="tsitpap8")
            cpu0 = time.perf_counter() - t0
            extras["config0_default_adaptive"] = {
                "workload": "configs[0]: N=1e4, ts=0:0.05:100, the reference default integrator_alg=TsitPap8() at tol=1e-12 -> DOP853 tableau under "
                            "OrdinaryDiffEq's controller for TsitPap8 (the tableau itself is not available offline)",
                "gpu_wall_s": wall0, "gpu_kernel_ms": i0["kernel_ms"], "gpu_attempted_steps": i0["steps_accepted"] + i0["steps_rejected"],
                "cpu_port_s": cpu0, "cpu_cores": thr, "cpu_attempted_steps": r0["steps_accepted"] + r0["steps_rejected"],
                "max_abs_diff_mean_jz": float(np.max(np.abs(r0["sums"][:, 0] / j / 10000 - g0))),
                "hermite_output": {"note": "output='hermite': the reference's interpolated outputs (src/TWA.jl:180) instead of landing on every ts[k]",
                                   "gpu_wall_s": wall0h, "gpu_kernel_ms": i0h["kernel_ms"], "gpu_attempted_steps": i0h["steps_accepted"] + i0h["steps_rejected"],
                                   "max_abs_diff_mean_jz_vs_stepping": float(np.max(np.abs(g0h - g0)))}}
            # the same dense grid in the throughput regime (the docs' larger ensembles): 1e6 trajectories
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for _ in range(2):
                    W0 = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
                    t0 = time.perf_counter()
                    TWA.average(system, observable=jz, N=1000000, ts=ts0, distribution=W0)
                    wall1 = time.perf_counter() - t0
            i1 = dict(TWA.last_info)
            extras["config0_grid_1e6_trajectories"] = {
                "workload": "configs[0]'s grid and integrator (ts=0:0.05:100, tol=1e-12: about one output per 1.5 attempted steps) with N=1e6",
                "gpu_wall_s": wall1, "gpu_kernel_ms": i1["kernel_ms"], "attempted_steps": i1["steps_accepted"] + i1["steps_rejected"],
                "attempted_steps_per_s": (i1["steps_accepted"] + i1["steps_rejected"]) / (i1["kernel_ms"] * 1e-3),
                "lane_efficiency": (i1["steps_accepted"] + i1["steps_rejected"]) / max(1, i1["lane_steps"])}
            # SURVEY 8f ranks 2-4 (the callers on the other side of `integrate`), small instances, for the record
            CSm, ESP = pkg.ClassicalSystems, pkg.EnergyShellProjections
            sysw = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
            ew = -1.35
            Qs = np.linspace(CD.minimum_nonnegative_Q_for_ϵ(sysw, ew), CD.maximum_Q_for_ϵ(sysw, ew), 128)
            Ps = np.linspace(0, CD.maximum_P_for_ϵ(sysw, ew), 128)
            pts = np.array([CD.Point(sysw, Q=Q, P=P, p=0, ϵ=ew) for Q in Qs for P in Ps if CD.minimum_ϵ_for(sysw, P=P, p=0, Q=Q) <= ew])
            for _ in range(2):
                lam, stl = CSm.lyapunov_spectrum_batch(sysw, pts, 100, tol=1e-8, λtol=5e-5, verbose=False)
            il = CSm.last_lyapunov_info
            extras["note_8f"] = ("lyapunov_map_128 / sections_p0_t1000 / shell_quadrature are small instances (12.6e3 orbits leave most of a "
                                 "B200 idle in the tail); full-size numbers: profiles/r1_lyapunov_map_256.json, r1_sections_and_shell_quadrature.jsonl")
            extras["lyapunov_map_128"] = {"orbits": int(len(pts)), "variational_steps": int(il["nsteps"].sum()), "kernel_ms": il["kernel_ms"],
                                          "steps_per_s": float(il["nsteps"].sum() / (il["kernel_ms"] * 1e-3)), "converged": int((stl == 0).sum())}
            for _ in range(2):
                t0 = time.perf_counter()
                sec = CSm.poincare_section_batch(sysw, pts, 1000.0, coef=[0, 0, 0, 1], tol=1e-10, max_events=512, packed=True)
                wall_s = time.perf_counter() - t0
            extras["sections_p0_t1000"] = {"orbits": int(len(pts)), "steps": int(sec["nsteps"].sum()), "crossings": int(len(sec["t"])),
                                           "kernel_ms": sec["kernel_ms"], "wall_s": wall_s,
                                           "steps_per_s": float(sec["nsteps"].sum() / (sec["kernel_ms"] * 1e-3)),
                                           "max_abs_p_on_section": float(np.abs(sec["u"][:, 3]).max())}
            for _ in range(2):
                Qm, Pm, mat = ESP.matrix_QP_iint_dqdp_delta(system, f=lambda x: [1, x[1] ** 2 + x[3] ** 2], ϵ=-0.5, res=0.01)
            ish = dict(ESP.last_info)
            extras["shell_quadrature_res0.01"] = {"cells": ish["cells"], "nodes": ish["nodes"], "kernel_ms": ish["kernel_ms"],
                                                  "nodes_per_s": ish["nodes"] / (ish["kernel_ms"] * 1e-3),
                                                  "volume_over_closed_form": float(np.nansum(mat[..., 0]) * 1e-4 / CD.energy_shell_volume(system, -0.5))}
        except Exception as e:   # never lose the headline line over the extras
            extras["error_single_gpu_extras"] = repr(e)

    if rank == 0:
        if not (abs(float(last[0]) + 1.0) < 0.02 and steps_done == args.steps * N_total * 14000):
            raise SystemExit(f"bench self-check failed: <Jz>/j(t=0) = {float(last[0])}, steps counted {steps_done} "
                             f"(expected {args.steps * N_total * 14000}): the per-rank arenas were not merged")
        nominal_tflops = 148 * 64 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
        line = {
            "metric": "FP64 trajectory-steps/s", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n_traj), "trajectories_total": N_total, "steps_per_trajectory": 14000,
                       "cache": "no input tensors (Philox sampling is fused into the kernel); a 256 MiB buffer is written before every step to flush the 126 MB L2",
                       "python_gc": "disabled inside the two timed loops (as timeit does), collected right before them",
                       "block_threads": args.block or 256, "integrator": "RK4", "sampler": "coherent_Wigner_HWxSU2, Philox4x32-10",
                       "reference_arm_note": "the CPU arms (cpu_baseline, --impl reference) integrate a bounded SAMPLE of this ensemble per step "
                                             "(same workload, fewer trajectories): the metric is a rate"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "step_timing": {"host_wall_ms": step_wall_ms, "library_total_ms": step_lib_ms,
                            "note": "rank 0, per timed step of the device-resident leg: host clock around the public call (L2 flush + TWA.average) and the "
                                    "library's own first-memset-to-last-copy time; ms_per_step is the CUDA-event time of the whole loop / steps"},
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops, "frac_nominal": achieved_tflops / nominal_tflops, "peak_nominal": nominal_tflops,
                         "k0_ms": k0_ms, "k0_flops": peak_tflops * 1e12 * k0_ms * 1e-3,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "flops_per_step": FLOPS_PER_STEP["dicke_rk4"],
                         "k0_shapes_tflops": {"8x256_threads_per_sm": k0a_tflops, "32x32_threads_per_sm": k0c_tflops},
                         "peak_source": "K0 DFMA-chain micro-benchmarks run in this process (dtwa_fp64_peak / dtwa_fp64_chains, the faster of two launch "
                                        "shapes: k0_flops in k0_ms, best of 4-5); MEASURED_PEAKS.json has "
                                        "no FP64 entry; peak_nominal = 148 SM x 64 DFMA/clk x 2 x sm_max_mhz",
                         "kernel": "dtwa_jit_ensemble (NVRTC specialisation of ensemble_kernel<DICKE,RK4> for this reducer set)", "kernel_ms_per_step": kernel_ms / args.steps,
                         "kernel_steps_per_s_per_gpu": kern_rate,
                         "fp64_pipe_frac_est": kern_rate * FP64_INSTR_PER_STEP["dicke_rk4"] / (nominal_tflops * 1e12 / 2.0),
                         "note": "frac is ALGORITHMIC flops (136/step) over the measured DFMA peak, frac_nominal over the nominal one; fp64_pipe_frac_est counts the 88 "
                                 "FP64-pipe instructions/step of the SASS hot loop (w = g = 1 specialisation; 96 in general: extras.general_parameters), each of "
                                 "which occupies one DFMA issue slot, over the nominal issue rate"},
            "cpu_baseline": cpu,
            "extras": extras,
            "result_check": {"mean_jz_t0": float(last[0]), "mean_jz_t5": float(last[100]), "mean_jz_t100": float(last[-1])},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
