#!/bin/bash
# round 2, GPU call 28: step-gap diagnostic with the flush timed separately and Python's collector frozen
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python tools/diag_step_gap.py 1e7 > gpurun_out/r2c28_step_gap.log 2>&1; echo "rc $?" | tee -a gpurun_out/r2c28_status.txt
