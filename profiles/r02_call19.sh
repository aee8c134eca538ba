#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ROUNDS="1" AB_FRAMES=24 bash profiles/ab_tc3.sh run cbias pp0 nopponly nomufunopp cbias 2>&1 | tee gpurun_out/r02_ablation2_tc3.txt
