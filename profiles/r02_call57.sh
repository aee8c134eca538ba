#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ROUNDS="1 2 3" AB_FRAMES=40 bash profiles/ab_tc3.sh run hint0 hint2k hint20k 2>&1 | tee gpurun_out/r02_ab_tc3_waithint.txt
