#!/bin/bash
# round 2, GPU call 42: 1-D histogram window flush with four bins per shared-memory load (padded staging rows)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="timeout 300 python tools/profile_case.py --state B --tstep 0.05 --tend 20 --n 4000000 --reps 3 --jit 1"
{
echo "== two 1-D-vs-time histograms"; $P --hist
echo "== docs chain of three"; $P --mode chain3
echo "== average"; $P
echo "== two 1-D-vs-time histograms, interpreter"; timeout 300 python tools/profile_case.py --state B --tstep 0.05 --tend 20 --n 4000000 --reps 2 --jit 0 --hist
} > $O/r2c42_hist.log 2>&1
echo "variants done" | tee -a $O/r2c42_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c42_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c42_status.txt
