#!/bin/bash
# round 2, GPU call 38: longer rings on the sparse grid of config 4 (lane efficiency against ring traffic)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
for w in 256 384 512; do for alg in TsitPap8 DP5; do echo "== $alg ring $w"; DTWA_WINDOW=$w $P --alg $alg; done; done
} > $O/r2c38_long_ring.log 2>&1
echo done | tee -a $O/r2c38_status.txt
