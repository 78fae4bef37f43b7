#!/bin/bash
# A/B bench lines (custom + val) for the libraries given as arguments, then the parity tests on the LAST one
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "$@"; do
  n=$(basename $lib .so)
  for mode in ${MODES:-custom val}; do
    NEXAR_LIB=$PWD/$lib timeout 120 python bench.py --workload ${WORKLOAD:-cfg2} --mode $mode --no-cpu-baseline --no-e2e > gpurun_out/ab_${n}_${mode}.log 2>&1
  done
done
[ -n "$NOTEST" ] || NEXAR_LIB=$PWD/${@: -1} timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
