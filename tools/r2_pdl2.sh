#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
for v in 0 2 0 2; do
  NEXAR_RESIZE_VARIANT=$v timeout 200 python bench.py --steps 200 --warmup 10 --mode val --no-cpu-baseline --no-e2e >> gpurun_out/bench_val_v$v.log 2>&1
done
NEXAR_RESIZE_VARIANT=0 timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e >> gpurun_out/bench_custom_v0.log 2>&1
true
