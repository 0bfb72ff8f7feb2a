#!/bin/bash
# round 2, GPU call 14: host-to-host latency of small calls; DP5 at 16 CTAs/SM (default now) next to DOP853
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
timeout 300 python tools/small_call_latency.py --jit 2
timeout 300 python tools/small_call_latency.py --jit 0
timeout 300 python tools/small_call_latency.py --jit 2 --alg RK4
} > $O/r2c14_small_call.log 2>&1; echo "small call rc $?" | tee -a $O/r2c14_status.txt
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
echo "== DP5 (16 CTAs/SM default)"; $P --alg DP5
echo "== DP5 bloch"; $P --alg DP5 --chart bloch
echo "== DP8"; $P --alg DP8
} > $O/r2c14_variants.log 2>&1
echo "variants done" | tee -a $O/r2c14_status.txt
