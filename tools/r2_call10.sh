#!/bin/bash
# round 2, GPU call 10: smoke, 16 CTAs/SM and literal coefficients for the adaptive kernels, CTAs/SM scaling, all five configs, ncu of the current DOP853 kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c10_smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/r2c10_status.txt
P="python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
echo "== TsitPap8 12 CTAs"; $P --alg TsitPap8
echo "== TsitPap8 12 CTAs literals"; DTWA_JIT_DEFINES="-DDTWA_COEF_LITERALS=1" $P --alg TsitPap8
echo "== TsitPap8 16 CTAs"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16" $P --alg TsitPap8
echo "== TsitPap8 16 CTAs literals"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16 -DDTWA_COEF_LITERALS=1" $P --alg TsitPap8
echo "== DP8 16 CTAs literals"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16 -DDTWA_COEF_LITERALS=1" $P --alg DP8
echo "== DP5 16 CTAs"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16" $P --alg DP5
echo "== DP5 20 CTAs"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=20" $P --alg DP5
echo "== TsitPap8 bloch 16 CTAs literals"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16 -DDTWA_COEF_LITERALS=1" $P --alg TsitPap8 --chart bloch
echo "== TsitPap8 8 CTAs/SM (of 12)"; $P --alg TsitPap8 --ctas-per-sm 8
echo "== TsitPap8 4 CTAs/SM (of 12)"; $P --alg TsitPap8 --ctas-per-sm 4
} > $O/r2c10_variants.log 2>&1
echo "variants done" | tee -a $O/r2c10_status.txt
timeout 600 python tools/run_configs.py > $O/r2_configs_one_gpu_share.jsonl 2> $O/r2c10_configs.err; echo "configs rc $?" | tee -a $O/r2c10_status.txt
C1="python tools/profile_case.py --alg TsitPap8 --state B --tstep 1 --tend 200 --n 568320 --reps 1 --jit 1"
$C1 > $O/r2c10_plain_dop853.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_dop853_cfg4 $C1 > $O/r2c10_ncu_dop853.log 2>&1
echo "ncu dop853 rc $?" | tee -a $O/r2c10_status.txt
