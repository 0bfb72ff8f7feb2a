#!/bin/bash
# round 2, GPU call 37 (2 GPUs): the multi-GPU tests of HEAD incl. the overlapped upload inside a two-device context
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_ensemble.py -m gpu -q -rs > $O/r2c37_pytest_2gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/r2c37_status.txt
tail -5 $O/r2c37_pytest_2gpu.log
