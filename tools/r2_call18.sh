#!/bin/bash
# round 2, GPU call 18: lean loop control of the DP5 specialised kernel (compile-time output mode, maxiters as a select, ts[k+1] prefetched)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/r2c18_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c18_status.txt
P="timeout 300 python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
Q="timeout 300 python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 1000000 --reps 2 --jit 1 --tol 1e-12"
{
echo "== DP5 cfg4"; $P --alg DP5
echo "== DP5 cfg4 hermite"; $P --alg DP5 --output hermite
echo "== DP5 cfg4 bloch"; $P --alg DP5 --chart bloch
echo "== DP8 cfg4"; $P --alg DP8
echo "== DP8 cfg4 hermite"; $P --alg DP8 --output hermite
echo "== DP5 dense (cfg0-like, 1e6)"; $Q --alg DP5
echo "== Tsit5 dense hermite"; $Q --alg Tsit5 --output hermite
echo "== TsitPap8 dense (cfg0-like, 1e6)"; $Q --alg TsitPap8
} > $O/r2c18_variants.log 2>&1
echo "variants done" | tee -a $O/r2c18_status.txt
timeout 300 python tools/small_call_latency.py --jit 2 --alg DP5 --tol 1e-8 > $O/r2c18_small_call_dp5.log 2>&1; echo "small call rc $?" | tee -a $O/r2c18_status.txt
