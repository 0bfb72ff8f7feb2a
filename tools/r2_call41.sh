#!/bin/bash
# round 2, GPU call 41: A/B on one box -- histogram window flush by rows against the flat index with an integer division, and the
# staging budget (output times per window) of the fixed-step kernels with histograms
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
P="timeout 300 python tools/profile_case.py --state B --tstep 0.05 --tend 20 --n 4000000 --reps 3 --jit 1"
{
echo "== hist rows"; $P --hist
echo "== hist flat"; DTWA_JIT_DEFINES="-DDTWA_FLUSH_FLAT=1" $P --hist
echo "== hist rows"; $P --hist
echo "== hist flat"; DTWA_JIT_DEFINES="-DDTWA_FLUSH_FLAT=1" $P --hist
for kb in 12 24 40 48; do echo "== hist rows budget $kb KB"; DTWA_WBUDGET_KB=$kb $P --hist; done
echo "== chain3 rows"; $P --mode chain3
echo "== chain3 flat"; DTWA_JIT_DEFINES="-DDTWA_FLUSH_FLAT=1" $P --mode chain3
} > $O/r2c41_hist_ab.log 2>&1
echo "done" | tee -a $O/r2c41_status.txt
