#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest60.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke60.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_v6.json 2> gpurun_out/r02_bench_n1_v6.err; echo "bench rc=$?"
