#!/bin/bash
# one full ncu capture each of K2 (colour) and K3 (geometry), custom mode, after the plain run exited 0
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:colour -s 4 -c 1 -f -o gpurun_out/prof_k2 $CMD > gpurun_out/ncu_k2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:geometry -s 4 -c 1 -f -o gpurun_out/prof_k3 $CMD > gpurun_out/ncu_k3.log 2>&1
