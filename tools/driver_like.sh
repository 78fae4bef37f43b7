#!/bin/bash
# what the driver runs at round end, with its usual small K / W
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/drv_reference.log 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/drv_bench.log 2>&1
true
