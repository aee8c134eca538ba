#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py tests/test_train_tc_gpu.py tests/test_cuda_golden.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest50.log
python profiles/prof_train_batched.py 2>&1 | tail -2 | tee gpurun_out/r02_train_prof50.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:'k_mp_edge_tc|k_gate_bwd|k_rows_gemm' --csv --log-file gpurun_out/r02_launches_train_v7.csv python profiles/prof_train_batched.py --no-graph > gpurun_out/ncu_t7.log 2>&1; echo "ncu rc=$?"
