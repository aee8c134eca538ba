#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest30.log
python profiles/prof_train_batched.py 2>&1 | tail -3 | tee gpurun_out/r02_train_prof30.log
