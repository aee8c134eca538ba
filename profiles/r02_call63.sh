#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="matches_oracle or workload_size or decision_exercising or small_weights or range_overflow or golden or two_feature"
TMPNN_LIB=build/lib_nofence.so timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_cuda_golden.py -m gpu -x -q -k "$K" 2>&1 | tail -3 | tee gpurun_out/r02_pytest63.log
AB_ROUNDS="1 2 3" AB_FRAMES=40 bash profiles/ab_tc3.sh run xfence nofence 2>&1 | tee gpurun_out/r02_ab_tc3_xfence.txt
