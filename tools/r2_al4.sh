#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3.log 2>&1
timeout 200 python bench.py --workload cfg3 --mode val --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3_val.log 2>&1
true
