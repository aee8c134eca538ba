#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_train_engine_gpu.py tests/test_engine_gpu.py -m gpu -q -k "flat or oracle_at_bench or no_tp or reinit" > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest3.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r02_bench_n2.err
