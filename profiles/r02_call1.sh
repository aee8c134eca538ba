#!/bin/bash
# round 2, first GPU call: tests, smoke, bench, launch list, DRAM counters, one full capture of the edge kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke1.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke1.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
SHORT="python bench.py --steps 1 --warmup 1 --frames 12 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
$SHORT > gpurun_out/r02_short.json 2> gpurun_out/r02_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/r02_launches_base.csv $SHORT > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_mp_edge_tc3|k_aggregate_dets|k_det_prepare' -s 60 -c 36 --csv --log-file gpurun_out/r02_dram_base.csv $SHORT > gpurun_out/ncu_d.log 2>&1
echo "ncu dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc3 -s 30 -c 1 -f -o gpurun_out/r02_tc3_base $SHORT > gpurun_out/ncu_f.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
