#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TMPNN_LIB=build/lib_h3.so timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest27.log
AB_ROUNDS="1 2" AB_FRAMES=24 bash profiles/ab_tc3.sh run v3 h3 2>&1 | tee gpurun_out/r02_ab_h3.txt
TMPNN_LIB=build/lib_h3_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -1
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_h3.npy
