#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TMPNN_LIB=build/lib_pipepstg.so timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest21.log
AB_ROUNDS="1 2" AB_FRAMES=24 bash profiles/ab_tc3.sh run cbias pipe pstg pipepstg 2>&1 | tee gpurun_out/r02_ab_pipe.txt
