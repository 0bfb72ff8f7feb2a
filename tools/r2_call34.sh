#!/bin/bash
# round 2, GPU call 34: blocking-sync waits against spinning waits for the long kernels: bench headline, alternating, on one box;
# small-call latency both ways
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for i in 1 2 3; do
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras > $O/r2c34_bench_block_$i.json 2> $O/r2c34_bench_block_$i.err; echo "bench block $i rc $?" | tee -a $O/r2c34_status.txt
DTWA_SPIN_WAIT=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras > $O/r2c34_bench_spin_$i.json 2> $O/r2c34_bench_spin_$i.err; echo "bench spin $i rc $?" | tee -a $O/r2c34_status.txt
done
{
timeout 300 python tools/small_call_latency.py --jit 1 --calls 30
DTWA_SPIN_WAIT=1 timeout 300 python tools/small_call_latency.py --jit 1 --calls 30
} 2>&1 | grep "^N=" > $O/r2c34_small_call.log; echo "small call done" | tee -a $O/r2c34_status.txt
