#!/bin/bash
# round 2, GPU call 32: late host calls and Python's collector -- 16 steps with / without gc in the diagnostic; bench headline x3 with
# the collector off inside the timed loops
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python tools/diag_step_gap.py 1e7 short > $O/r2c32_step_gap.log 2>&1; echo "diag rc $?" | tee -a $O/r2c32_status.txt
for i in 1 2 3; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > $O/r2c32_bench_$i.json 2> $O/r2c32_bench_$i.err; echo "bench $i rc $?" | tee -a $O/r2c32_status.txt; done
