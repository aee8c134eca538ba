#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_engine_gpu.py tests/test_train_tc_gpu.py tests/test_cuda_golden.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest34.log
python profiles/prof_train_batched.py 2>&1 | tail -3 | tee gpurun_out/r02_train_prof34.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_v5.csv python profiles/prof_train_batched.py --no-graph > gpurun_out/ncu_t5.log 2>&1; echo "ncu train rc=$?"
SHORT="python bench.py --steps 1 --warmup 1 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 130 --csv --log-file gpurun_out/r02_launches_infer_v3.csv $SHORT > gpurun_out/ncu_l3.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_mp_edge_tc3|k_aggregate_blocks|k_det_prepare' -s 120 -c 48 --csv --log-file gpurun_out/r02_dram_v3.csv $SHORT > gpurun_out/ncu_d3.log 2>&1
echo "ncu dram rc=$?"
