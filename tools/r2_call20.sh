#!/bin/bash
# round 2, GPU call 20: validation of HEAD (GPU tests) + where the hardware places one-warp CTAs (%smid / %warpid of every warp
# stream of the adaptive kernel) + the adaptive kernels with 2 / 4 warp streams per CTA (DTWA_ADAPT_WARPS)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2c20_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c20_status.txt
D="timeout 120 python tools/profile_case.py --state B --tstep 1 --tend 20 --n 56832 --reps 1 --jit 1 --alg DP8"
DTWA_JIT_DEFINES="-DDTWA_DIAG_WARPID" $D > $O/r2c20_slots_w1.log 2>&1
DTWA_ADAPT_WARPS=4 DTWA_JIT_DEFINES="-DDTWA_DIAG_WARPID" $D > $O/r2c20_slots_w4.log 2>&1
echo "slots done" | tee -a $O/r2c20_status.txt
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
for alg in TsitPap8 DP8 DP5; do
echo "== $alg 1 warp/CTA"; $P --alg $alg
echo "== $alg 4 warps/CTA"; DTWA_ADAPT_WARPS=4 $P --alg $alg
echo "== $alg 2 warps/CTA"; DTWA_ADAPT_WARPS=2 $P --alg $alg
done
} > $O/r2c20_variants.log 2>&1
echo "variants done" | tee -a $O/r2c20_status.txt
