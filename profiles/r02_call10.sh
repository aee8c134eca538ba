#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_tc_gpu.py tests/test_train_engine_gpu.py tests/test_cuda_golden.py -m gpu -q > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest10.log
TR="python profiles/prof_train_batched.py"
$TR > gpurun_out/r02_train_prof10.log 2>&1; tail -3 gpurun_out/r02_train_prof10.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train10.csv $TR --no-graph > gpurun_out/ncu_t.log 2>&1
echo "ncu train launches rc=$?"
