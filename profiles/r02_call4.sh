#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_tc_gpu.py -m gpu -q -x > gpurun_out/r02_pytest4a.log 2>&1; echo "tc tests rc=$?"; tail -25 gpurun_out/r02_pytest4a.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest4.log
timeout 300 python profiles/prof_train_batched.py > gpurun_out/r02_train_prof4.log 2>&1; echo "prof rc=$?"; grep -E "build s|host enqueue|step ms" gpurun_out/r02_train_prof4.log
