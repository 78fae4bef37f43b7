#!/bin/bash
# parity tests, A/B bench lines for the libraries given as arguments, then one full ncu capture of K2 + K3 (custom mode)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./tools/ab.sh "$@"
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"colour|geometry" -s 6 -c 2 -o gpurun_out/prof_k23 $CMD > gpurun_out/ncu_k23.log 2>&1
