#!/bin/bash
# A/B with repeats: every library twice, interleaved, each run appended to its own log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
for lib in "$@"; do
  n=$(basename $lib .so)
  NEXAR_LIB=$PWD/$lib timeout 120 python bench.py --steps 100 --warmup 10 --mode ${MODE:-custom} $EXTRA --no-cpu-baseline --no-e2e >> gpurun_out/ab_${n}.log 2>&1
done
done
true
