#!/bin/bash
# round 2, GPU call 5: GPU tests; software-pipelined adaptive hot loop vs the classic one; Hermite outputs after the tlim fix; ncu
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/r2c5_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c5_status.txt
P="python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
NP="-DDTWA_NO_PIPELINED_LOOP=1"
{
echo "== TsitPap8 pipelined"; $P --alg TsitPap8
echo "== TsitPap8 classic"; DTWA_JIT_DEFINES="$NP" $P --alg TsitPap8
echo "== DP8 pipelined"; $P --alg DP8
echo "== DP8 classic"; DTWA_JIT_DEFINES="$NP" $P --alg DP8
echo "== DP5 pipelined"; $P --alg DP5
echo "== DP5 classic"; DTWA_JIT_DEFINES="$NP" $P --alg DP5
echo "== TsitPap8 bloch pipelined"; $P --alg TsitPap8 --chart bloch
echo "== TsitPap8 bloch classic"; DTWA_JIT_DEFINES="$NP" $P --alg TsitPap8 --chart bloch
echo "== TsitPap8 pipelined 16 CTAs/SM"; DTWA_JIT_DEFINES="-DDTWA_ADAPT_MIN_BLOCKS=16" $P --alg TsitPap8
echo "== TsitPap8 hermite"; $P --alg TsitPap8 --output hermite
echo "== TsitPap8 bloch hermite"; $P --alg TsitPap8 --output hermite --chart bloch
Q="python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 1000000 --reps 2 --jit 1 --tol 1e-12"
echo "== cfg0-like step pipelined"; $Q --alg TsitPap8
echo "== cfg0-like step classic"; DTWA_JIT_DEFINES="$NP" $Q --alg TsitPap8
echo "== cfg0-like hermite"; $Q --alg TsitPap8 --output hermite
echo "== cfg0-like bloch step"; $Q --alg TsitPap8 --chart bloch
} > $O/r2c5_variants.log 2>&1
echo "variants done" | tee -a $O/r2c5_status.txt
C1="python tools/profile_case.py --alg TsitPap8 --state B --tstep 1 --tend 200 --n 568320 --reps 1 --jit 1"
$C1 > $O/r2c5_plain_dop853.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dtwa_jit_ensemble -c 1 -f -o $O/r2_dop853_cfg4_v8 $C1 > $O/r2c5_ncu_dop853.log 2>&1
echo "ncu dop853 rc $?" | tee -a $O/r2c5_status.txt
