#!/bin/bash
# Same-box A/B of compile-time variants of the re-staged edge kernel.
#   here:        bash profiles/ab_tc3.sh build "NAME:-DFLAG=1 -DOTHER=2" ...   -> build/lib_NAME.so
#   on the box:  bash profiles/ab_tc3.sh run NAME ...                           (3 alternating rounds)
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p build
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -I include $flags \
      trackmpnn_b200/csrc/{mp_step,mp_step_tc,mp_step_tc3,graph_index,graph_ops,train,train_tc,hungarian,gat,features}.cu -o build/lib_$name.so 2>/dev/null
    echo built build/lib_$name.so "($flags)"
  done
else
  for i in ${AB_ROUNDS:-1 2 3}; do for v in "$@"; do
    echo -n "$v "
    TMPNN_LIB=build/lib_$v.so timeout 100 python bench.py --steps 2 --warmup 2 --frames ${AB_FRAMES:-40} --skip-cpu --skip-e2e --skip-train --skip-c4 --skip-check 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,4), round(d['roofline']['avg_launch_ms'],4), 'edge GB/s', round(d['roofline']['achieved'],1), 'agg ms', round(d['roofline_aggregation']['avg_launch_ms'],4), 'agg frac', round(d['roofline_aggregation']['frac'],3))"
  done; done
fi
