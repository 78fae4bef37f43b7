#!/bin/bash
# parity tests, then custom-mode bench lines with (variant 0) and without (variant 2) the programmatic overlap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
for v in 0 2 0 2; do
  NEXAR_RESIZE_VARIANT=$v timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e >> gpurun_out/bench_pdl_v$v.log 2>&1
done
NEXAR_RESIZE_VARIANT=0 timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3_v0.log 2>&1
NEXAR_RESIZE_VARIANT=2 timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3_v2.log 2>&1
true
