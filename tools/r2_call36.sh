#!/bin/bash
# round 2, GPU call 36: ts[k+1] prefetched for every pair and streaming (evict-first) ring accesses, against the previous forms,
# on the dense grid of configs[0] (10^6 trajectories) and on config 4; GPU tests with the new forms
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
PD="timeout 300 python tools/profile_case.py --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 1000000 --reps 2 --jit 1"
PS="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
for alg in TsitPap8 DP5; do
for d in "" "-DDTWA_PREFETCH_TS=0" "-DDTWA_RING_STREAMING=0" "-DDTWA_PREFETCH_TS=0 -DDTWA_RING_STREAMING=0"; do
echo "== dense $alg [$d]"; DTWA_JIT_DEFINES="$d" $PD --alg $alg
echo "== sparse $alg [$d]"; DTWA_JIT_DEFINES="$d" $PS --alg $alg
done; done
echo "== dense TsitPap8 ring 256 []"; DTWA_WINDOW=256 $PD --alg TsitPap8
echo "== dense TsitPap8 ring 128 []"; DTWA_WINDOW=128 $PD --alg TsitPap8
echo "== dense TsitPap8 variance []"; $PD --alg TsitPap8 --mode variance
echo "== dense TsitPap8 hist []"; $PD --alg TsitPap8 --hist
} > $O/r2c36_variants.log 2>&1
echo "variants done" | tee -a $O/r2c36_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c36_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c36_status.txt
