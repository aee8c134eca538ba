#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest5.log
timeout 300 python profiles/prof_train_batched.py > gpurun_out/r02_train_prof5.log 2>&1; echo "prof rc=$?"; grep -E "build s|host enqueue|step ms" gpurun_out/r02_train_prof5.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train5.csv python profiles/prof_train_batched.py > gpurun_out/ncu_t5.log 2>&1; echo "ncu rc=$?"
