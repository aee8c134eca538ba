#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest47.log
SHORT="python bench.py --steps 2 --warmup 2 --frames 40 --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check"
sel='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e9,4), "edge ms", round(d["roofline"]["avg_launch_ms"],4), "agg ms", round(d["roofline_aggregation"]["avg_launch_ms"],4), "ms/step", round(d["ms_per_step"],1), d["phases"]["update_ms_per_tick"], d["phases"]["forward_ms_per_tick"], d["phases"]["decode_ms_per_tick"], d["gpu_launches"])'
for i in 1 2; do timeout 200 $SHORT 2>/dev/null | python -c "$sel"; done 2>&1 | tee gpurun_out/r02_short47.txt
