#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SHORT="python bench.py --steps 1 --warmup 1 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 140 --csv --log-file gpurun_out/r02_launches_infer_v4.csv $SHORT > gpurun_out/ncu_l4.log 2>&1
echo "ncu launches rc=$?"
