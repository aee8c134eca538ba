#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 ./profiles/micro/gather4_test 2>&1 | tee gpurun_out/r02_micro_gather4.txt
K="matches_oracle or workload_size or decision_exercising or small_weights or range_overflow or golden or two_feature"
TMPNN_LIB=build/lib_tma.so timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_cuda_golden.py -m gpu -x -q -k "$K" 2>&1 | tail -5 | tee gpurun_out/r02_pytest59.log
AB_ROUNDS="1 2 3" AB_FRAMES=40 bash profiles/ab_tc3.sh run cpasync tma 2>&1 | tee gpurun_out/r02_ab_tc3_tma.txt
