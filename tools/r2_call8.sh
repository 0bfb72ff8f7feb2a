#!/bin/bash
# round 2, GPU call 8: batched transposed ring flush; K0c sweep (chains x warps)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/r2c8_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c8_status.txt
timeout 300 python tools/k0_chains.py > $O/r2_k0_chains.log 2>&1; echo "k0 chains rc $?" | tee -a $O/r2c8_status.txt
P="python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 1"
{
echo "== TsitPap8"; $P --alg TsitPap8
echo "== DP8"; $P --alg DP8
echo "== DP5"; $P --alg DP5
echo "== TsitPap8 bloch"; $P --alg TsitPap8 --chart bloch
echo "== TsitPap8 hermite"; $P --alg TsitPap8 --output hermite
echo "== TsitPap8 survival"; $P --alg TsitPap8 --mode survival
echo "== TsitPap8 interpreter (jit 0)"; python tools/profile_case.py --state B --tstep 1 --tend 1000 --n 1000000 --reps 2 --jit 0 --alg TsitPap8
Q="python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 1000000 --reps 2 --jit 1 --tol 1e-12"
echo "== cfg0-like step"; $Q --alg TsitPap8
echo "== cfg0-like hermite"; $Q --alg TsitPap8 --output hermite
echo "== cfg0-like bloch hermite"; $Q --alg TsitPap8 --output hermite --chart bloch
echo "== cfg0-like step 1e4 traj"; python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 10000 --reps 3 --jit 1 --tol 1e-12 --alg TsitPap8
echo "== cfg0-like hermite 1e4 traj"; python tools/profile_case.py --state A --tstep 0.05 --tend 100 --n 10000 --reps 3 --jit 1 --tol 1e-12 --alg TsitPap8 --output hermite
} > $O/r2c8_variants.log 2>&1
echo "variants done" | tee -a $O/r2c8_status.txt
