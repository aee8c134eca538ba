#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_v2.json 2> gpurun_out/r02_bench_n1_v2.err; echo "bench rc=$?"
SHORT="python bench.py --steps 1 --warmup 1 --frames 12 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/r02_launches_infer_v2.csv $SHORT > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_mp_edge_tc3|k_aggregate_dets|k_det_prepare' -s 60 -c 36 --csv --log-file gpurun_out/r02_dram_v2.csv $SHORT > gpurun_out/ncu_d.log 2>&1
echo "ncu dram rc=$?"
