#!/bin/bash
# round 2, GPU call 19: DOP853 with the lean loop control and literal coefficients (no constant loads for ptxas to hoist)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
echo "== DP8 baseline"; $P --alg DP8
echo "== DP8 lean + literals"; DTWA_JIT_DEFINES="-DDTWA_LEAN_DOP853=1 -DDTWA_COEF_LITERALS=1" $P --alg DP8
echo "== DP8 lean"; DTWA_JIT_DEFINES="-DDTWA_LEAN_DOP853=1" $P --alg DP8
echo "== TsitPap8 lean + literals"; DTWA_JIT_DEFINES="-DDTWA_LEAN_DOP853=1 -DDTWA_COEF_LITERALS=1" $P --alg TsitPap8
echo "== DP5 literals"; DTWA_JIT_DEFINES="-DDTWA_COEF_LITERALS=1" $P --alg DP5
} > $O/r2c19_variants.log 2>&1
echo "variants done" | tee -a $O/r2c19_status.txt
{
timeout 300 python tools/small_call_latency.py --jit 1 --calls 20
DTWA_JIT_DEFINES="-DDTWA_LEAN_DOP853=1 -DDTWA_COEF_LITERALS=1" timeout 300 python tools/small_call_latency.py --jit 1 --calls 20
} 2>&1 | grep "^N=" > $O/r2c19_small_call.log; echo "small call done" | tee -a $O/r2c19_status.txt
