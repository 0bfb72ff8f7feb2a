#!/bin/bash
# round 2, GPU call 43: as many warp streams as give every lane the same number of trajectories (DTWA_BALANCE_STREAMS), on one rank's
# share of the strong-scaling config (10^6 trajectories: 17.6 per lane) and on 2 x 10^6
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
for n in 1000000 2000000 500000; do
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n $n --reps 2 --jit 1"
for alg in TsitPap8 DP5; do
echo "== $alg N=$n full grid"; $P --alg $alg
echo "== $alg N=$n balanced"; DTWA_BALANCE_STREAMS=1 $P --alg $alg
done; done
} > $O/r2c43_balance.log 2>&1
echo done | tee -a $O/r2c43_status.txt
