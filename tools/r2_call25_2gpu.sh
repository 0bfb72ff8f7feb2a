#!/bin/bash
# round 2, GPU call 25 (2 GPUs): the multi-GPU tests and bench.py at N=2 for HEAD (Tsit5, ring-length rule, NVML clock sampler);
# a headline-only N=1 run to see the sampler's cost (ms_per_step against kernel_ms_per_step)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2c25_gpus.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > $O/r2c25_bench_n1_headline.json 2> $O/r2c25_bench_n1.err; echo "bench n1 rc $?" | tee -a $O/r2c25_status.txt
timeout 1200 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_ensemble.py -m gpu -q -rs > $O/r2c25_pytest_2gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c25_status.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2c25_bench_n2.json 2> $O/r2c25_bench_n2.err; echo "bench n2 rc $?" | tee -a $O/r2c25_status.txt
tail -5 $O/r2c25_bench_n2.err
head -3 $O/clocks_rank0.csv
