#!/bin/bash
# band-count sweep of the fast resize kernel (NEXAR_FAST_BANDS) with the programmatic overlap in place
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for nb in 0 2 3 4 5 6 8; do
  NEXAR_FAST_BANDS=$nb timeout 120 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_bands$nb.log 2>&1
done
for nb in 0 4 8 12 16; do
  NEXAR_FAST_BANDS=$nb timeout 120 python bench.py --workload cfg2s --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_s_bands$nb.log 2>&1
done
for nb in 0 4; do
  NEXAR_FAST_BANDS=$nb timeout 120 python bench.py --mode val --steps 100 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_val_bands$nb.log 2>&1
done
true
