#!/bin/bash
# round 2, GPU call 16 (8 GPUs): the multi-GPU tests and bench.py at N=2, 4 and 8 exactly as the driver launches it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2c16_gpus.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_ensemble.py -m gpu -q -rs > $O/r2c16_pytest_multi_gpu.log 2>&1; echo "pytest multi-gpu rc $?" | tee -a $O/r2c16_status.txt
for n in 8 4 2; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 > $O/r2c16_bench_n$n.json 2> $O/r2c16_bench_n$n.err; echo "bench n$n rc $?" | tee -a $O/r2c16_status.txt
done
