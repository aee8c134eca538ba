#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest45.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_smoke45.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_v4.json 2> gpurun_out/r02_bench_n1_v4.err; echo "bench rc=$?"
SHORT="python bench.py --steps 1 --warmup 1 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
ncu --set full --clock-control none --import-source on -k regex:'k_mp_edge_tc3' -s 60 -c 1 -f -o gpurun_out/r02_mp_edge_tc3_final $SHORT > gpurun_out/ncu_f3.log 2>&1
echo "ncu full tc3 rc=$?"
