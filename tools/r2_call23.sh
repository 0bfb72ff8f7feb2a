#!/bin/bash
# round 2, GPU call 23: the controller's lg2(err) from the factors of err (parallel to the square root) and the loop vote as two
# VOTE.ANY -- against the previous forms (-DDTWA_LE_FROM_ERR, -DDTWA_VOTE_REDUX), config 4; GPU tests with the new forms
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
for alg in TsitPap8 DP8 DP5; do
echo "== $alg new"; $P --alg $alg
echo "== $alg old le"; DTWA_JIT_DEFINES="-DDTWA_LE_FROM_ERR=1" $P --alg $alg
echo "== $alg old vote"; DTWA_JIT_DEFINES="-DDTWA_VOTE_REDUX=1" $P --alg $alg
echo "== $alg old both"; DTWA_JIT_DEFINES="-DDTWA_LE_FROM_ERR=1 -DDTWA_VOTE_REDUX=1" $P --alg $alg
done
} > $O/r2c23_variants.log 2>&1
echo "variants done" | tee -a $O/r2c23_status.txt
{
timeout 300 python tools/small_call_latency.py --jit 1 --calls 20
DTWA_JIT_DEFINES="-DDTWA_LE_FROM_ERR=1 -DDTWA_VOTE_REDUX=1" timeout 300 python tools/small_call_latency.py --jit 1 --calls 20
} 2>&1 | grep "^N=" > $O/r2c23_small_call.log; echo "small call done" | tee -a $O/r2c23_status.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2c23_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c23_status.txt
