#!/bin/bash
# round 2, GPU call 39: ncu --set full (with source) of the RK4 kernel with the two-histogram chain of configs[2]
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
C="python tools/profile_case.py --state B --tstep 0.05 --tend 20 --n 2000000 --reps 1 --jit 1 --hist"
timeout 300 $C > $O/r2c39_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_rk4_hist_chain $C > $O/r2c39_ncu.log 2>&1
echo "ncu rc $?" | tee -a $O/r2c39_status.txt
cat $O/r2c39_plain.log | tail -2
