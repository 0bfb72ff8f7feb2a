#!/bin/bash
# round 2, GPU call 29 (8 GPUs): bench.py at N=8 for HEAD, exactly as the driver launches it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2c29_gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2c29_bench_n8.json 2> $O/r2c29_bench_n8.err; echo "bench n8 rc $?" | tee -a $O/r2c29_status.txt
tail -3 $O/r2c29_bench_n8.err
