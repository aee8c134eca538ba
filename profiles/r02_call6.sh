#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python profiles/graph_train_step.py > gpurun_out/r02_graph_train6.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02_graph_train6.log
timeout 600 python -m pytest tests/test_train_engine_gpu.py tests/test_cuda_golden.py -m gpu -q > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest6.log
