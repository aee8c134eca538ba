#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest40.log
AB_ROUNDS="1 2" AB_FRAMES=40 bash profiles/ab_tc3.sh run prep4 prep8 2>&1 | tee gpurun_out/r02_ab_prep.txt
SHORT="python bench.py --steps 1 --warmup 1 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_block_fill|k_det_prepare' -s 150 -c 6 --csv --log-file gpurun_out/r02_launches_fill.csv $SHORT > gpurun_out/ncu_l5.log 2>&1
grep -c . gpurun_out/r02_launches_fill.csv
