#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ROUNDS="1 2 3" AB_FRAMES=40 bash profiles/ab_tc3.sh run ld0 ld2 2>&1 | tee gpurun_out/r02_ab_tc3_ld2.txt
