#!/bin/bash
# round 2, GPU call 2: GPU tests on the new build, adaptive-kernel variants (v6), Bloch chart, 2-D histograms, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c2_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c2_status.txt
P="python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
echo "== v6 TsitPap8 12 CTAs/SM"; $P --alg TsitPap8
echo "== v6 TsitPap8 16 CTAs/SM (128 regs)"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16" $P --alg TsitPap8
echo "== v6 DP8 12"; $P --alg DP8
echo "== v6 DP8 scipy 12"; $P --alg DP8 --controller scipy
echo "== v6 TsitPap8 bloch 12"; $P --alg TsitPap8 --chart bloch
echo "== v6 TsitPap8 bloch 16"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16" $P --alg TsitPap8 --chart bloch
echo "== v6 DP5 12"; $P --alg DP5
echo "== v6 DP5 16"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16" $P --alg DP5
echo "== v6 DP5 bloch 12"; $P --alg DP5 --chart bloch
echo "== v6 TsitPap8 general w g"; $P --alg TsitPap8 --w 0.8 --g 0.66
R="python tools/profile_case.py --n 4000000 --tend 20 --reps 2 --jit 1"
echo "== RK4 average"; $R
echo "== RK4 average bloch"; $R --chart bloch
echo "== RK4 average bloch general"; $R --chart bloch --w 0.8 --g 0.66
echo "== RK4 hist2d docs size"; $R --mode hist2d
echo "== RK4 hist2d 81x81 staged"; $R --mode hist2d --small-bins
echo "== RK4 hist2d 81x81 direct"; DTWA_NO_STAGE2D=1 $R --mode hist2d --small-bins
echo "== RK4 animate"; $R --mode animate
echo "== RK4 chain (two 1-D)"; $R --hist
echo "== RK4 chain3"; $R --mode chain3
echo "== RK4 variance"; $R --mode variance
echo "== RK4 survival"; $R --mode survival
} > $O/r2c2_variants.log 2>&1
echo "variants done" | tee -a $O/r2c2_status.txt
C0="python -c \"import __graft_entry__ as g; p=g.load_package(); c=p.default_context(); print(c.fp64_peak())\""
bash -c "$C0" > $O/r2c2_plain_k0.log 2>&1 && ncu --metrics sm__warps_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,sm__maximum_warps_per_active_cycle_pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:fp64_peak -c 2 --csv --log-file $O/r2_k0_occupancy.csv bash -c "$C0" > $O/r2c2_ncu_k0.log 2>&1
echo "ncu k0 rc $?" | tee -a $O/r2c2_status.txt
C1="python tools/profile_case.py --alg TsitPap8 --state B --tstep 1 --tend 200 --n 568320 --reps 1 --jit 1"
$C1 > $O/r2c2_plain_dop853.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_dop853_cfg4_v6 $C1 > $O/r2c2_ncu_dop853.log 2>&1
echo "ncu dop853 rc $?" | tee -a $O/r2c2_status.txt
C2="python tools/profile_case.py --n 2000000 --tend 10 --reps 1 --mode hist2d --jit 1"
$C2 > $O/r2c2_plain_hist2d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_hist2d_v2 $C2 > $O/r2c2_ncu_hist2d.log 2>&1
echo "ncu hist2d rc $?" | tee -a $O/r2c2_status.txt
C3="python tools/profile_case.py --n 2000000 --tend 10 --reps 1 --jit 1 --chart bloch"
$C3 > $O/r2c2_plain_rk4bloch.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_rk4_bloch $C3 > $O/r2c2_ncu_rk4bloch.log 2>&1
echo "ncu rk4 bloch rc $?" | tee -a $O/r2c2_status.txt
ls -la $O/*.ncu-rep | tee -a $O/r2c2_status.txt
