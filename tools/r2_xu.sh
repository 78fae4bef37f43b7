#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
for lib in exp_xu0 exp_xu1 exp_xu2; do
  NEXAR_LIB=$PWD/vision_collision_detection_b200/$lib.so timeout 120 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e >> gpurun_out/ab_$lib.log 2>&1
done
done
for lib in exp_xu1 exp_xu2; do
NEXAR_LIB=$PWD/vision_collision_detection_b200/$lib.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_$lib.log 2>&1; echo pytest=$? >> gpurun_out/pytest_$lib.log
done
true
