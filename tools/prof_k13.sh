#!/bin/bash
# parity tests, then one full ncu capture each of K1 and K3 (custom mode)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"resize_fast|geometry" -s 6 -c 2 -o gpurun_out/prof_k13 $CMD > gpurun_out/ncu_k13.log 2>&1
