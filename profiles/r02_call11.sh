#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest11.log
python profiles/prof_train_batched.py > gpurun_out/r02_train_prof11.log 2>&1; tail -2 gpurun_out/r02_train_prof11.log
