#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ROUNDS="1 2" bash profiles/ab_tc3.sh run base smr80 2>&1 | tee gpurun_out/r02_ab2.txt
TMPNN_LIB=build/lib_smr80.so timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "workload or matches_oracle" 2>&1 | tail -2
