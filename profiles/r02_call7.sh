#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest7.log
timeout 300 python profiles/graph_train_step.py > gpurun_out/r02_graph_train7.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02_graph_train7.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench7.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_dram'], d['roofline_aggregation']['frac'], d['roofline_aggregation']['frac_dram'])
print(d['train']); print(d['c4']['value'] if 'value' in d['c4'] else d['c4'])
PY
