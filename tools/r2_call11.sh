#!/bin/bash
# round 2, GPU call 11: K0c sweep with 1/2/3 register operands; GPU tests; bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python tools/k0_chains.py > $O/r2_k0_chains.log 2>&1; echo "k0 chains rc $?" | tee -a $O/r2c11_status.txt
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/r2c11_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c11_status.txt
timeout 1200 python bench.py > $O/r2c11_bench_n1.json 2> $O/r2c11_bench_n1.err; echo "bench rc $?" | tee -a $O/r2c11_status.txt
