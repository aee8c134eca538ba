#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest52.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke52.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_v5.json 2> gpurun_out/r02_bench_n1_v5.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_v8.csv python profiles/prof_train_batched.py --no-graph > gpurun_out/ncu_t8.log 2>&1; echo "ncu train rc=$?"
