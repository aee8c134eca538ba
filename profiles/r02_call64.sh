#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="matches_oracle or workload_size or decision_exercising or small_weights or range_overflow or golden or two_feature"
TMPNN_LIB=build/lib_xs1.so timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_cuda_golden.py -m gpu -x -q -k "$K" 2>&1 | tail -3 | tee gpurun_out/r02_pytest64.log
AB_ROUNDS="1 2 3" AB_FRAMES=40 bash profiles/ab_tc3.sh run xs0 xs1 2>&1 | tee gpurun_out/r02_ab_tc3_xsplit.txt
TMPNN_LIB=build/lib_xs1_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -1
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_xsplit.npy
