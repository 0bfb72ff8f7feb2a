#!/bin/bash
# round 2, GPU call 33: final validation of HEAD as the driver runs it -- smoke(), pytest -m gpu, bench.py --impl reference, bench.py
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2c33_smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/r2c33_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c33_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c33_status.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c33_bench_reference.json 2> $O/r2c33_bench_reference.err; echo "reference arm rc $?" | tee -a $O/r2c33_status.txt
timeout 1200 python bench.py --steps 20 --warmup 3 > $O/r2c33_bench_n1.json 2> $O/r2c33_bench_n1.err; echo "bench rc $?" | tee -a $O/r2c33_status.txt
