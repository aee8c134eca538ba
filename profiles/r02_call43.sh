#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2_v3.json 2> gpurun_out/r02_bench_n2_v3.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r02_bench_n2_v3.err
