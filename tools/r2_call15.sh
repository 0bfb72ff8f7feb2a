#!/bin/bash
# round 2, GPU call 15: validation of HEAD -- GPU tests, smoke, the default bench line and the reference arm, small-call latency, launch list of the bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/r2c15_pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c15_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c15_smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/r2c15_status.txt
timeout 1200 python bench.py > $O/r2c15_bench_n1.json 2> $O/r2c15_bench_n1.err; echo "bench rc $?" | tee -a $O/r2c15_status.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c15_bench_ref.json 2> $O/r2c15_bench_ref.err; echo "bench ref rc $?" | tee -a $O/r2c15_status.txt
timeout 300 python tools/small_call_latency.py --jit 2 > $O/r2c15_small_call.log 2>&1; echo "small call rc $?" | tee -a $O/r2c15_status.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-e2e"
timeout 600 $B > $O/r2c15_plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_steps2_warmup3_final.csv $B > $O/r2c15_ncu_launches.log 2>&1
echo "ncu launches rc $?" | tee -a $O/r2c15_status.txt
