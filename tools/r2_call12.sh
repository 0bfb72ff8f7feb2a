#!/bin/bash
# round 2, GPU call 12: K0d sweep (other-pipe instructions woven into the DFMA chains)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python tools/k0_mix.py > $O/r2_k0_mix.log 2>&1; echo "k0 mix rc $?" | tee -a $O/r2c12_status.txt
