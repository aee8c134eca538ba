#!/bin/bash
# in-place transpose (h images) + far-endpoint MMAs first: parity, A/B against the round's base, role timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_cuda_golden.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_pytest15.log
AB_ROUNDS="1 2" bash profiles/ab_tc3.sh run base inplace 2>&1 | tee gpurun_out/r02_ab_inplace.txt
TMPNN_LIB=build/libtmpnn_trace.so timeout 200 python profiles/trace_tc.py run 2>&1 | tail -2
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_inplace.npy
