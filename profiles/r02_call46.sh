#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_n8_v2.json 2> gpurun_out/r02_bench_n8_v2.err; echo "bench n8 rc=$?"
tail -c 300 gpurun_out/r02_bench_n8_v2.err
