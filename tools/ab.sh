#!/bin/bash
# A/B: parity tests on the in-tree build, then bench lines for each library given as argument (paths relative to the repo)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
for lib in "$@"; do
  n=$(basename $lib .so)
  for mode in ${MODES:-custom val}; do
    NEXAR_LIB=$PWD/$lib timeout 300 python bench.py --mode $mode --no-cpu-baseline --no-e2e > gpurun_out/ab_${n}_${mode}.log 2>&1
  done
done
