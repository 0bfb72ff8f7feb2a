#!/bin/bash
# round 2, GPU call 30: the overlapped upload of host-supplied initial conditions (split launch) -- GPU tests, bench headline with
# and without it, and the ncu launch list of bench.py for HEAD
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c30_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c30_status.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > $O/r2c30_bench_overlap.json 2> $O/r2c30_bench_overlap.err; echo "bench overlap rc $?" | tee -a $O/r2c30_status.txt
DTWA_NO_COPY_OVERLAP=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > $O/r2c30_bench_nooverlap.json 2> $O/r2c30_bench_nooverlap.err; echo "bench no-overlap rc $?" | tee -a $O/r2c30_status.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c30_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $O/r2c30_ncu.log 2>&1; echo "ncu rc $?" | tee -a $O/r2c30_status.txt
