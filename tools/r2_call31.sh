#!/bin/bash
# round 2, GPU call 31: step-gap diagnostic of the host-array path with and without the overlapped upload; pinned H2D rate alone
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python tools/diag_step_gap.py 1e7 > gpurun_out/r2c31_step_gap.log 2>&1; echo "rc $?" | tee -a gpurun_out/r2c31_status.txt
