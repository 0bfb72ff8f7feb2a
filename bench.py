#!/usr/bin/env python
"""bench.py -- the headline measurement: FP64 trajectory-steps/s of the TWA ensemble on B200.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun launches N ranks)
    python bench.py --impl reference --gpus N --steps K ...  (the restated reference CPU path)

Workload (BASELINE.json configs[1]): Dicke w = w0 = 1, gamma = 2 gamma_c = 1,
j = 100, coherent state of the normal-phase ground state x0 = [0,0,0,0] (the quench), 10^7
trajectories PER GPU, ts = 0:0.05:100 (2001 output times), fixed-step RK4 with dt = 2^-7 (7 equal
sub-steps per output interval => 14000 steps per trajectory), observable Weyl.Jz(j)/j averaged at
every output time.  One bench "step" = one whole ensemble.

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
            f"ts=0:0.05:100 (nt=2001), RK4 dt=2^-7, <Jz>/j")


def clocks_sampler_start(path):
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    TWA, CD = pkg.TWA, pkg.ClassicalDicke
    stream = torch.cuda.current_stream()
    ctx = pkg.Context([local], stream=stream.cuda_stream)
    pkg.set_default_context(ctx)
    if args.block or args.ctas_per_sm:
        ctx.set_launch(args.block, args.ctas_per_sm)
    if world > 1:
        pkg.set_distributed(rank, world, pkg.distributed.torch_allreduce(local))

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

    # ---- roofline denominator: K0 DFMA chains ----
    peak_tflops, _ = ctx.fp64_peak()

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
    sampler = clocks_sampler_start(clk_path) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    steps_done, kernel_ms, launches, last = 0, 0.0, 0, None
    for _ in range(args.steps):
        last, info = step_device()
        steps_done += info["steps_accepted"]       # all-reduced: whole job
        kernel_ms += info["kernel_ms"]
        launches += info["launches"]
    e1.record(stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    if sampler:
        sampler.terminate()
        sampler.wait()
    t = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms = float(t[0]), float(t[1])
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

        for _ in range(max(1, args.warmup - 2)):
            step_e2e()
        barrier()
        e0.record(stream)
        st, e2e_out = 0, None
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_out, info = step_e2e()
            st += info["steps_accepted"]
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": st / (float(t[0]) * 1e-3), "unit": "trajectory-steps/s",
               "h2d_bytes_per_step": int(n_loc * 4 * 8 * world + len(ts) * 8 * world), "d2h_bytes_per_step": int(len(ts) * 8 * world + len(ts) * 8 * world),
               "steps": n_e2e, "note": "TWA.average with distribution=samples_from_array(pinned host u0[N][4]); H2D of the initial conditions and D2H of the per-time sums inside the timed region"}
        # same samples, same kernel: the two legs must agree to rounding
        e2e["max_abs_diff_vs_device_leg"] = float(np.max(np.abs(e2e_out - last))) if world == 1 else None
        del host

    # ---- CPU baseline (rank 0, N = 1 only): the restated reference CPU path on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
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

    # ---- not the headline: the adaptive path (BASELINE.json configs[3]) on one GPU, for the record ----
    extras = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            Bpt = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5)
            tsa = TWA.jrange(0, 1, 1000)
            extras = {}
            for alg in ("DP8", "DP5"):
                for n_ad in (20000, 1000000):   # (the small call warms the specialisation up)
                    Wa = TWA.coherent_Wigner_HWxSU2(Bpt, j=j, seed=1)
                    TWA.average(system, observable=jz, N=n_ad, ts=tsa, distribution=Wa, integrator_alg=alg, tol=1e-8)
                ia = TWA.last_info
                st_ad = ia["steps_accepted"] + ia["steps_rejected"]
                extras["adaptive_" + alg.lower()] = {
                    "workload": "configs[3]: chaotic state, tol=1e-8, ts=0:1:1000, 1e6 trajectories (one GPU's share of 8e6)",
                    "attempted_steps_per_s": st_ad / (ia["kernel_ms"] * 1e-3), "kernel_ms": ia["kernel_ms"],
                    "lane_efficiency": st_ad / max(1, ia["lane_steps"]), "rejected_fraction": ia["steps_rejected"] / max(1, st_ad),
                    "specialised": ia["jit"]}
            # BASELINE.json configs[0] (the CPU-runnable case): N = 1e4, adaptive 8th order at the reference's default
            # tol = 1e-12, GPU next to the oracle port on all host cores, same Philox samples
            from oracle import oracle as O
            ts0 = TWA.jrange(0, WORK["t_step"], WORK["t_end"])
            for _ in range(2):
                W0 = TWA.coherent_Wigner_HWxSU2(WORK["center"], j=j, seed=WORK["seed"])
                t0 = time.perf_counter()
                g0 = TWA.average(system, observable=jz, N=10000, ts=ts0, distribution=W0)
                wall0 = time.perf_counter() - t0
            i0 = dict(TWA.last_info)
            sp0 = O.make_sampler(O.SAMPLER_HWXSU2, WORK["center"], 1.0 / j, seed=WORK["seed"])
            cj = float(np.sqrt(j * (j + 1)))
            thr = host_threads()
            t0 = time.perf_counter()
            r0 = O.ensemble(O.SYS_DICKE, WORK["par"], sp0, 10000, np.asarray(ts0), values=[(2, 0, cj)], alg=O.ALG_DOP853, tol=1e-12,
                            workers=thr, fast=True)
            cpu0 = time.perf_counter() - t0
            extras["config0_default_adaptive"] = {
                "workload": "configs[0]: N=1e4, ts=0:0.05:100, DOP853 tol=1e-12 (TsitPap8 is not available offline)",
                "gpu_wall_s": wall0, "gpu_kernel_ms": i0["kernel_ms"], "gpu_attempted_steps": i0["steps_accepted"] + i0["steps_rejected"],
                "cpu_port_s": cpu0, "cpu_cores": thr, "cpu_attempted_steps": r0["steps_accepted"] + r0["steps_rejected"],
                "max_abs_diff_mean_jz": float(np.max(np.abs(r0["sums"][:, 0] / j / 10000 - g0)))}
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
            extras = dict(extras or {}, error=repr(e))

    if rank == 0:
        if not (abs(float(last[0]) + 1.0) < 0.02 and steps_done == args.steps * N_total * 14000):
            raise SystemExit(f"bench self-check failed: <Jz>/j(t=0) = {float(last[0])}, steps counted {steps_done} "
                             f"(expected {args.steps * N_total * 14000}): the per-rank arenas were not merged")
        line = {
            "metric": "FP64 trajectory-steps/s", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n_traj), "trajectories_total": N_total, "steps_per_trajectory": 14000,
                       "cache": "no input tensors (Philox sampling is fused into the kernel); a 256 MiB buffer is written before every step to flush the 126 MB L2",
                       "block_threads": args.block or 256, "integrator": "RK4", "sampler": "coherent_Wigner_HWxSU2, Philox4x32-10"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops, "traffic": traffic, "traffic_source": traffic_src,
                         "flops_per_step": FLOPS_PER_STEP["dicke_rk4"],
                         "peak_source": "K0 DFMA-chain micro-benchmark run in this process (dtwa_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; nominal 37.2 TFLOP/s at 1965 MHz",
                         "kernel": "dtwa_jit_ensemble (NVRTC specialisation of ensemble_kernel<DICKE,RK4> for this reducer set)", "kernel_ms_per_step": kernel_ms / args.steps,
                         "kernel_steps_per_s_per_gpu": kern_rate,
                         "fp64_pipe_frac_est": kern_rate * FP64_INSTR_PER_STEP["dicke_rk4"] / (peak_tflops * 1e12 / 2.0),
                         "note": "frac is ALGORITHMIC flops (136/step) over the measured DFMA peak; fp64_pipe_frac_est counts the 88 FP64-pipe instructions/step of the SASS hot loop (w = g = 1 specialisation), each of which occupies one DFMA issue slot"},
            "cpu_baseline": cpu,
            "extras": extras,
            "result_check": {"mean_jz_t0": float(last[0]), "mean_jz_t5": float(last[100]), "mean_jz_t100": float(last[-1])},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
