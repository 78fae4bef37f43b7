#!/bin/bash
# cfg2 with 4 / 5 / 6 / 7 bands per frame (resize kernel grid = bands x frames)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in ${BANDS:-5 6 4 7 5 6}; do
  for mode in ${MODES:-custom val}; do
    NEXAR_FAST_BANDS=$b timeout 120 python bench.py --workload ${WORKLOAD:-cfg2} --mode $mode --no-cpu-baseline --no-e2e >> gpurun_out/bands_${b}_${mode}.log 2>&1
  done
done
