#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_cuda_golden.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest18.log
AB_ROUNDS="1 2" AB_FRAMES=24 bash profiles/ab_tc3.sh run inplace cbias cbiaspf 2>&1 | tee gpurun_out/r02_ab_cbias.txt
