#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TMPNN_LIB=build/lib_both.so timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest23.log
AB_ROUNDS="1 2" AB_FRAMES=24 bash profiles/ab_tc3.sh run v2 preload headmbar both 2>&1 | tee gpurun_out/r02_ab_preload.txt
