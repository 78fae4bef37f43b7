#!/bin/bash
# one full ncu capture of the kernel matching $1 (regex) in mode ${2:-custom}; skip count ${3:-4}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --mode ${2:-custom} ${4} --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-4} -c 1 -f -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
