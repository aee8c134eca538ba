#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest8.log
timeout 600 python bench.py --steps 3 --warmup 3 --skip-train --skip-c4 --skip-cpu > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench8.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'])
print('edge', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'])
print('agg', d['roofline_aggregation']['frac'], d['roofline_aggregation']['frac_dram'], d['roofline_aggregation']['avg_launch_ms'], d['roofline_aggregation']['share_of_step'])
print(d['phases'])
PY
