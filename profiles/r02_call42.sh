#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest42.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_v3.json 2> gpurun_out/r02_bench_n1_v3.err; echo "bench rc=$?"
SHORT="python bench.py --steps 1 --warmup 1 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 140 --csv --log-file gpurun_out/r02_launches_infer_v5.csv $SHORT > gpurun_out/ncu_l7.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_mp_edge_tc3|k_aggregate_blocks|k_det_prepare' -s 150 -c 50 --csv --log-file gpurun_out/r02_dram_v4.csv $SHORT > gpurun_out/ncu_d4.log 2>&1
echo "ncu dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'^k_aggregate_blocks$' -s 30 -c 1 -f -o gpurun_out/r02_aggregate_blocks $SHORT > gpurun_out/ncu_f2.log 2>&1
echo "ncu full agg rc=$?"
