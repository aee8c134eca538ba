#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TMPNN_LIB=build/libtmpnn_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -2
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_inplace.npy
