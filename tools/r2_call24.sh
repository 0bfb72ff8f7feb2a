#!/bin/bash
# round 2, GPU call 24: Tsit5 (the real tableau) -- GPU tests, throughput next to DP5 on config 4 and on configs[0]'s grid; bench.py
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c24_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c24_status.txt
P="timeout 300 python tools/profile_case.py --reps 2 --jit 1"
{
for alg in Tsit5 DP5; do
echo "== $alg config 4 (state B, tol 1e-8, ts=0:1:1000, 1e6)"; $P --state B --tstep 1 --tend 1000 --n 1000000 --alg $alg
echo "== $alg cfg0 grid (state A, tol 1e-12, ts=0:0.05:100, 2e5)"; $P --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 200000 --alg $alg
done
echo "== TsitPap8 cfg0 grid 1e6 (automatic ring)"; $P --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 1000000 --alg TsitPap8
} > $O/r2c24_tsit5.log 2>&1
echo "variants done" | tee -a $O/r2c24_status.txt
timeout 900 python bench.py > $O/r2c24_bench_n1.json 2> $O/r2c24_bench_n1.err; echo "bench rc $?" | tee -a $O/r2c24_status.txt
