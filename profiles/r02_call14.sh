#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ROUNDS="1 2" bash profiles/ab_tc3.sh run base u8 na u8na g16 2>&1 | tee gpurun_out/r02_ab3_agg.txt
