#!/bin/bash
# round 2, GPU call 22: ring length of the adaptive kernels against the density of the output grid (outputs per attempted step)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
PC="timeout 300 python tools/profile_case.py --reps 2 --jit 1"
{
# dense grid of configs[0] (tol 1e-12, ts = 0:0.05:100), 10^6 trajectories
for w in 48 64 96; do echo "== cfg0 average ring $w"; DTWA_WINDOW=$w $PC --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 1000000 --alg TsitPap8; done
for w in 64 96 128; do echo "== cfg0 variance ring $w"; DTWA_WINDOW=$w $PC --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 1000000 --alg TsitPap8 --mode variance; done
for w in 64 256; do echo "== cfg0 hist ring $w"; DTWA_WINDOW=$w $PC --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 1000000 --alg TsitPap8 --hist; done
# state B, tol 1e-8 (h ~ 0.09), t <= 200: spacing 1 / 0.5 / 0.25 / 0.1
for dt in 1 0.5 0.25 0.1; do for w in 64 128 256; do echo "== cfg4 spacing $dt ring $w"; DTWA_WINDOW=$w $PC --state B --tstep $dt --tend 200 --tol 1e-8 --n 1000000 --alg TsitPap8; done; done
# fewer trajectories: 2 x 10^5 (3.5 per lane) on the dense grid
for w in 64 256; do echo "== cfg0 average 2e5 ring $w"; DTWA_WINDOW=$w $PC --state A --tstep 0.05 --tend 100 --tol 1e-12 --n 200000 --alg TsitPap8; done
} > $O/r2c22_ring_density.log 2>&1
echo "done" | tee -a $O/r2c22_status.txt
