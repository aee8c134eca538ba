#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_rows_gemm_tc' -c 2 -f -o gpurun_out/r02_rows_gemm_tc_final python profiles/prof_train_batched.py --no-graph > gpurun_out/ncu_f4.log 2>&1; echo "ncu rc=$?"
