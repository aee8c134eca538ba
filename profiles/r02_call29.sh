#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TMPNN_LIB=build/lib_n192_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -1
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_n192.npy
SHORT="python bench.py --steps 1 --warmup 1 --frames 12 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc3 -s 30 -c 1 -f -o gpurun_out/r02_mp_edge_tc3_v4 $SHORT > gpurun_out/ncu_f.log 2>&1
echo "ncu full tc3 rc=$?"
