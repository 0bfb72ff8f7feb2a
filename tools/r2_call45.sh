#!/bin/bash
# round 2, GPU call 45: K0d with IMAD (integer multiply-add on the FMA pipe -- where ptxas also sends half of its register moves)
# woven into the DFMA chains, next to IADD (ALU pipe) and FFMA
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python tools/k0_mix.py iadd ffma imad > gpurun_out/r2c45_k0_mix_imad.log 2>&1; echo "rc $?" | tee -a gpurun_out/r2c45_status.txt
