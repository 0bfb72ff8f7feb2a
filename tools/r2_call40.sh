#!/bin/bash
# round 2, GPU call 40: the 1-D histogram window flush without the integer division (two-histogram chain of configs[2], the docs'
# chain of three, one 2-D histogram for reference) + GPU tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="timeout 300 python tools/profile_case.py --state B --tstep 0.05 --tend 20 --n 4000000 --reps 2 --jit 1"
{
echo "== two 1-D-vs-time histograms"; $P --hist
echo "== docs chain of three"; $P --mode chain3
echo "== one 2-D histogram"; $P --mode hist2d
echo "== average"; $P
} > $O/r2c40_hist.log 2>&1
echo "variants done" | tee -a $O/r2c40_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c40_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c40_status.txt
