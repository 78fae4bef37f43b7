#!/bin/bash
# chunk-size sweep (NEXAR_CHUNK_CLIPS) at cfg2 and cfg3, custom mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
for c in 0 8 16; do
  NEXAR_CHUNK_CLIPS=$c timeout 120 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg2_chunk$c.log 2>&1
done
for c in 0 3 4 6 8 256; do
  NEXAR_CHUNK_CLIPS=$c timeout 200 python bench.py --steps 10 --warmup 3 --workload cfg3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3_chunk$c.log 2>&1
done
true
