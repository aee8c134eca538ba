#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_train_engine_gpu.py tests/test_train_tc_gpu.py tests/test_cuda_golden.py -m gpu -x -q -k "block_aggregation or matches_oracle or deferred or wide or train or golden or structured or edge_case or reinit" 2>&1 | tail -5 | tee gpurun_out/r02_pytest33.log
SHORT="python bench.py --steps 2 --warmup 2 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
sel='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e9,4), "edge ms", round(d["roofline"]["avg_launch_ms"],4), "agg ms", round(d["roofline_aggregation"]["avg_launch_ms"],4), "agg frac", round(d["roofline_aggregation"]["frac"],3), "ms/step", round(d["ms_per_step"],1))'
for i in 1 2; do
  echo -n "lists  "; timeout 200 $SHORT --list-aggregation 2>/dev/null | python -c "$sel"
  echo -n "blocks "; timeout 200 $SHORT 2>/dev/null | python -c "$sel"
done 2>&1 | tee gpurun_out/r02_ab_agg_blocks_v2.txt
C4="python bench.py --steps 2 --warmup 2 --win 20 --dets 200 --seqs-per-gpu 4 --frames 26 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
echo -n "c4 lists  "; timeout 300 $C4 --list-aggregation 2>/dev/null | python -c "$sel" | tee -a gpurun_out/r02_ab_agg_blocks_v2.txt
echo -n "c4 blocks "; timeout 300 $C4 2>/dev/null | python -c "$sel" | tee -a gpurun_out/r02_ab_agg_blocks_v2.txt
python profiles/prof_train_batched.py 2>&1 | tail -3 | tee gpurun_out/r02_train_prof33.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_v4.csv python profiles/prof_train_batched.py --no-graph > gpurun_out/ncu_t4.log 2>&1; echo "ncu train rc=$?"
