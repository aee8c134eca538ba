#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ROUNDS="1 2" AB_FRAMES=24 bash profiles/ab_tc3.sh run preload token 2>&1 | tee gpurun_out/r02_ab_token.txt
TMPNN_LIB=build/lib_token_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -1
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_token.npy
