#!/bin/bash
# ablation map of the edge kernel (results are WRONG by construction; only the time per launch matters)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_ROUNDS="1" AB_FRAMES=24 bash profiles/ab_tc3.sh run inplace nomufu noxmma nohmma noxcopy nopp noldtm noprev notrans nostg noown nommacopy noepimem inplace 2>&1 | tee gpurun_out/r02_ablation_tc3.txt
