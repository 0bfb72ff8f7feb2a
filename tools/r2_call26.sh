#!/bin/bash
# round 2, GPU call 26: where the 1-19 ms between a bench step and its kernel go (tools/diag_step_gap.py)
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python tools/diag_step_gap.py 1e7 > gpurun_out/r2c26_step_gap.log 2>&1; echo "rc $?" | tee -a gpurun_out/r2c26_status.txt
