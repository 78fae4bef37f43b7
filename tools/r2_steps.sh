#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 20 100 300 1000 20 300; do
  timeout 200 python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-e2e >> gpurun_out/bench_steps.log 2>&1
done
true
