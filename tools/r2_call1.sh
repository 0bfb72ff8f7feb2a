#!/bin/bash
# round 2, GPU call 1: GPU tests, one short bench line, adaptive-kernel variants, ncu captures (DOP853, 2-D histograms)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/r2c1_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2c1_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c1_status.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/r2c1_bench_n1.json 2> $O/r2c1_bench_n1.err; echo "bench rc $?" | tee -a $O/r2c1_status.txt
P="python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
echo "== DP8 scipy"; $P --alg DP8 --controller scipy
echo "== DP8 ode gains"; $P --alg DP8
echo "== TsitPap8 (generic PI gains)"; $P --alg TsitPap8
echo "== TsitPap8 window 128"; DTWA_WINDOW=128 $P --alg TsitPap8
echo "== TsitPap8 window 64"; DTWA_WINDOW=64 $P --alg TsitPap8
echo "== TsitPap8 128 regs (16 warps/SM)"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MAX_THREADS=128 -DDTWA_ADAPT_MIN_BLOCKS=4" $P --alg TsitPap8
echo "== TsitPap8 102 regs (20 warps/SM)"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MAX_THREADS=128 -DDTWA_ADAPT_MIN_BLOCKS=5" $P --alg TsitPap8
echo "== TsitPap8 general parameters"; $P --alg TsitPap8 --w 0.8 --g 0.66
echo "== DP5 ode gains"; $P --alg DP5
echo "== DP5 scipy"; $P --alg DP5 --controller scipy
echo "== RK4 average"; python tools/profile_case.py --n 4000000 --tend 20 --reps 2
echo "== RK4 hist2d"; python tools/profile_case.py --n 4000000 --tend 20 --reps 2 --mode hist2d
echo "== RK4 animate"; python tools/profile_case.py --n 4000000 --tend 20 --reps 2 --mode animate
echo "== RK4 chain (two 1-D)"; python tools/profile_case.py --n 4000000 --tend 20 --reps 2 --hist
echo "== RK4 chain3"; python tools/profile_case.py --n 4000000 --tend 20 --reps 2 --mode chain3
} > $O/r2c1_variants.log 2>&1
echo "variants done" | tee -a $O/r2c1_status.txt
C1="python tools/profile_case.py --alg TsitPap8 --state B --tstep 1 --tend 200 --n 568320 --reps 1 --jit 1"
$C1 > $O/r2c1_plain_dop853.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_dop853_cfg4 $C1 > $O/r2c1_ncu_dop853.log 2>&1
echo "ncu dop853 rc $?" | tee -a $O/r2c1_status.txt
C2="python tools/profile_case.py --n 2000000 --tend 10 --reps 1 --mode hist2d --jit 1"
$C2 > $O/r2c1_plain_hist2d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_hist2d_base $C2 > $O/r2c1_ncu_hist2d.log 2>&1
echo "ncu hist2d rc $?" | tee -a $O/r2c1_status.txt
C3="python tools/profile_case.py --n 200000 --tend 40 --tstep 0.1 --reps 1 --mode animate --jit 1"
$C3 > $O/r2c1_plain_animate.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_animate_base $C3 > $O/r2c1_ncu_animate.log 2>&1
echo "ncu animate rc $?" | tee -a $O/r2c1_status.txt
ls -la $O/*.ncu-rep | tee -a $O/r2c1_status.txt
