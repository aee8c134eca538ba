#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest9.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
SHORT="python bench.py --steps 1 --warmup 1 --frames 12 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
$SHORT > gpurun_out/r02_short.json 2> gpurun_out/r02_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/r02_launches_infer.csv $SHORT > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_mp_edge_tc3|k_aggregate_dets|k_det_prepare' -s 60 -c 36 --csv --log-file gpurun_out/r02_dram.csv $SHORT > gpurun_out/ncu_d.log 2>&1
echo "ncu dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_mp_edge_tc3 -s 30 -c 1 -f -o gpurun_out/r02_mp_edge_tc3 $SHORT > gpurun_out/ncu_f.log 2>&1
echo "ncu full tc3 rc=$?"
TR="python profiles/prof_train_batched.py --no-graph"
$TR > gpurun_out/r02_train_prof.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_batched.csv $TR > gpurun_out/ncu_t.log 2>&1
echo "ncu train launches rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_rows_gemm_tc -s 12 -c 1 -f -o gpurun_out/r02_rows_gemm_tc $TR > gpurun_out/ncu_g.log 2>&1
echo "ncu full gemm rc=$?"
python profiles/prof_train_batched.py > gpurun_out/r02_train_prof_graph.log 2>&1; tail -2 gpurun_out/r02_train_prof_graph.log
ls -la gpurun_out | tail -20
