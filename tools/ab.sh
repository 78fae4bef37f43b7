#!/bin/bash
# A/B of experiment libraries (NEXAR_LIB): one bench line per library; MODE / EXTRA select the bench variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "$@"; do
  n=$(basename $lib .so)
  NEXAR_LIB=$PWD/$lib timeout 120 python bench.py --steps 100 --warmup 10 --mode ${MODE:-custom} $EXTRA --no-cpu-baseline --no-e2e > gpurun_out/ab_${n}.log 2>&1
done
true
