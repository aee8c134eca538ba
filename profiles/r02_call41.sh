#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_chain_gpu.py tests/test_cuda_golden.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest41.log
SHORT="python bench.py --steps 1 --warmup 1 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_block_fill|k_det_prepare' -s 150 -c 6 --csv --log-file gpurun_out/r02_launches_fill2.csv $SHORT > gpurun_out/ncu_l6.log 2>&1
python profiles/prof_train_batched.py 2>&1 | tail -2 | tee gpurun_out/r02_train_prof41.log
