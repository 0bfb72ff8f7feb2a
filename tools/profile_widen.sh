#!/bin/bash
# ncu --set full captures of the SURVEY 8f kernels (one launch each), small sizes so ~40 replays stay cheap.
# usage (on the GPU box): bash tools/profile_widen.sh     -> gpurun_out/prof_r1_{lyapunov,events,shell}.ncu-rep
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:lyapunov_kernel -s 1 -c 1 -o gpurun_out/prof_r1_lyapunov python tools/bench_lyapunov.py --n 128 --cpu-points 0 > gpurun_out/ncu_lyap.log 2>&1
timeout 300 $NCU -k regex:events_kernel -s 1 -c 1 -o gpurun_out/prof_r1_events python tools/bench_widen.py --n 20000 --res 0.05 > gpurun_out/ncu_events.log 2>&1
timeout 300 $NCU -k regex:shell_quadrature_kernel -s 1 -c 1 -o gpurun_out/prof_r1_shell python tools/bench_widen.py --n 1024 --res 0.01 > gpurun_out/ncu_shell.log 2>&1
ls -la gpurun_out/*.ncu-rep
