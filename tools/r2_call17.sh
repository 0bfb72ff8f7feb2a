#!/bin/bash
# round 2, GPU call 17: ncu --set full of the DP5 specialised kernel at HEAD (16 CTAs/SM) on config 4
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
C1="python tools/profile_case.py --alg DP5 --state B --tstep 1 --tend 200 --n 757760 --reps 1 --jit 1"
timeout 300 $C1 > $O/r2c17_plain_dp5.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_dp5_cfg4 $C1 > $O/r2c17_ncu_dp5.log 2>&1
echo "ncu dp5 rc $?" | tee -a $O/r2c17_status.txt
ls -la $O/r2_dp5_cfg4.ncu-rep | tee -a $O/r2c17_status.txt
