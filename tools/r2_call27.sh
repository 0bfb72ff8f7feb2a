#!/bin/bash
# round 2, GPU call 27: bench.py headline twice and the step-gap diagnostic on the SAME box (per-step host times now in the line)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for i in 1 2; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > $O/r2c27_bench_$i.json 2> $O/r2c27_bench_$i.err; echo "bench $i rc $?" | tee -a $O/r2c27_status.txt; done
timeout 600 python tools/diag_step_gap.py 1e7 > $O/r2c27_step_gap.log 2>&1; echo "diag rc $?" | tee -a $O/r2c27_status.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --no-e2e > $O/r2c27_bench_3.json 2> $O/r2c27_bench_3.err; echo "bench 3 rc $?" | tee -a $O/r2c27_status.txt
