#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TMPNN_LIB=build/libtmpnn_trace.so timeout 300 python profiles/trace_tc.py run 2>&1 | tail -1
cp gpurun_out/tc_trace.npy gpurun_out/r02_tc_trace_tma.npy
SHORT="python bench.py --steps 1 --warmup 1 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 140 --csv --log-file gpurun_out/r02_launches_infer_v6.csv $SHORT > gpurun_out/ncu_l8.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_mp_edge_tc3|k_aggregate_blocks|k_det_prepare' -s 150 -c 50 --csv --log-file gpurun_out/r02_dram_v5.csv $SHORT > gpurun_out/ncu_d5.log 2>&1
echo "ncu dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_mp_edge_tc3' -s 60 -c 1 -f -o gpurun_out/r02_mp_edge_tc3_final2 $SHORT > gpurun_out/ncu_f5.log 2>&1
echo "ncu full tc3 rc=$?"
