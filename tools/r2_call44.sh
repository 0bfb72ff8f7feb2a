#!/bin/bash
# round 2, GPU call 44: last check of HEAD -- smoke(), pytest -m gpu, a short bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2c44_smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/r2c44_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c44_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c44_status.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2c44_bench_n1.json 2> $O/r2c44_bench_n1.err; echo "bench rc $?" | tee -a $O/r2c44_status.txt
