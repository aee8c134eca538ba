#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TMPNN_LIB=build/lib_oneteam.so timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest22.log
AB_ROUNDS="1 2" AB_FRAMES=24 bash profiles/ab_tc3.sh run base v2 oneteam 2>&1 | tee gpurun_out/r02_ab_oneteam.txt
TMPNN_LIB=build/lib_oneteam_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -1
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_oneteam.npy
TMPNN_LIB=build/lib_v2_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -1
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_v2.npy
