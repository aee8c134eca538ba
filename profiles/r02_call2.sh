#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest2.log
AB_ROUNDS="1 2" bash profiles/ab_tc3.sh run base sleep32 sleep200 nochk ks all 2>&1 | tee gpurun_out/r02_ab1.txt
