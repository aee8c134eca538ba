#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SHORT="python bench.py --steps 1 --warmup 1 --frames 12 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
$SHORT > gpurun_out/r02_short.json 2> gpurun_out/r02_short.err && \
ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc3 -s 30 -c 1 -f -o gpurun_out/r02_mp_edge_tc3_v3 $SHORT > gpurun_out/ncu_f.log 2>&1
echo "ncu full tc3 rc=$?"
