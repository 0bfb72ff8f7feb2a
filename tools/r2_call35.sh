#!/bin/bash
# round 2, GPU call 35: ncu --set full of (a) the Tsit5 specialised kernel on config 4, (b)/(c) the DOP853 kernel on configs[0]'s dense
# grid with the ring of 64 (now the default there) and of 256 output times
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
C1="python tools/profile_case.py --alg Tsit5 --state B --tstep 1 --tend 200 --n 757760 --reps 1 --jit 1"
C2="python tools/profile_case.py --alg TsitPap8 --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 200000 --reps 1 --jit 1"
timeout 300 $C1 > $O/r2c35_plain_tsit5.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_tsit5_cfg4 $C1 > $O/r2c35_ncu_tsit5.log 2>&1
echo "ncu tsit5 rc $?" | tee -a $O/r2c35_status.txt
timeout 300 $C2 > $O/r2c35_plain_dense64.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_dop853_dense_ring64 $C2 > $O/r2c35_ncu_dense64.log 2>&1
echo "ncu dense64 rc $?" | tee -a $O/r2c35_status.txt
DTWA_WINDOW=256 timeout 300 $C2 > $O/r2c35_plain_dense256.log 2>&1 && DTWA_WINDOW=256 timeout 900 ncu --set full --clock-control none -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_dop853_dense_ring256 $C2 > $O/r2c35_ncu_dense256.log 2>&1
echo "ncu dense256 rc $?" | tee -a $O/r2c35_status.txt
ls -la $O/*.ncu-rep | tee -a $O/r2c35_status.txt
