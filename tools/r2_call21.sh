#!/bin/bash
# round 2, GPU call 21: dense outputs on the adaptive path (configs[0]'s grid ts=0:0.05:100 at tol 1e-12, 10^6 trajectories):
# does a ring that stays in L2 (DTWA_WINDOW = 128 / 64 / 32 / 16 output times instead of 256) pay when almost every attempt emits?
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="timeout 300 python tools/profile_case.py --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 1000000 --reps 2 --jit 1 --alg TsitPap8"
{
for w in 256 128 64 32 16; do echo "== ring $w"; DTWA_WINDOW=$w $P; done
echo "== ring 256, variance of 4"; $P --mode variance
echo "== ring 32, variance of 4"; DTWA_WINDOW=32 $P --mode variance
echo "== ring 256, DP5"; timeout 300 python tools/profile_case.py --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 200000 --reps 2 --jit 1 --alg DP5
echo "== ring 32, DP5"; DTWA_WINDOW=32 timeout 300 python tools/profile_case.py --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 200000 --reps 2 --jit 1 --alg DP5
} > $O/r2c21_dense_ring.log 2>&1
echo "done" | tee -a $O/r2c21_status.txt
