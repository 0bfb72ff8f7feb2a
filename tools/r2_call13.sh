#!/bin/bash
# round 2, GPU call 13: two warp streams per warp for the adaptive kernels (DTWA_ADAPT_PAIR=1): GPU tests with it on, config-4 timings on/off
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
DTWA_ADAPT_PAIR=1 timeout 900 python -m pytest tests -m gpu -q -rs -x > $O/r2c13_pytest_gpu_pair.log 2>&1; echo "pytest(pair) rc $?" | tee -a $O/r2c13_status.txt
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
for alg in TsitPap8 DP8 DP5; do
echo "== $alg one stream"; DTWA_ADAPT_PAIR=0 $P --alg $alg
echo "== $alg pair"; DTWA_ADAPT_PAIR=1 $P --alg $alg
done
echo "== TsitPap8 pair literals"; DTWA_ADAPT_PAIR=1 DTWA_JIT_DEFINES="-DDTWA_COEF_LITERALS=1" $P --alg TsitPap8
echo "== TsitPap8 pair bloch"; DTWA_ADAPT_PAIR=1 $P --alg TsitPap8 --chart bloch
echo "== TsitPap8 one stream, docs spacing tol 1e-12 (config 0 at 1e6)"; DTWA_ADAPT_PAIR=0 timeout 300 python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 1000000 --reps 2 --jit 1 --alg TsitPap8 --tol 1e-12
echo "== TsitPap8 pair, docs spacing tol 1e-12"; DTWA_ADAPT_PAIR=1 timeout 300 python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 1000000 --reps 2 --jit 1 --alg TsitPap8 --tol 1e-12
} > $O/r2c13_variants.log 2>&1
echo "variants done" | tee -a $O/r2c13_status.txt
