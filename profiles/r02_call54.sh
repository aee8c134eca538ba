#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py tests/test_train_tc_gpu.py tests/test_cuda_golden.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest54.log
python profiles/prof_train_batched.py 2>&1 | tail -2 | tee gpurun_out/r02_train_prof54.log
