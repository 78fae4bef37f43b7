#!/bin/bash
# ncu evidence for the round (run under gpurun, one GPU): launch list of one bench command + one
# `--set full` capture of the dominant kernel.  The plain command must exit 0 first.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resize|colour|geometry|blur" -s 12 -c 16 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resize_fast -s 3 -c 1 -o gpurun_out/prof_k1_custom $CMD > gpurun_out/ncu_full.log 2>&1
python bench.py --workload cfg3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3.log 2>&1
