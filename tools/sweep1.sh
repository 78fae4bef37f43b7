#!/bin/bash
# experiment sweep (run under gpurun): variants of the fast resize kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --mode val --no-cpu-baseline --no-e2e"
echo "default"; $B 2>&1 | tail -1
for v in build/*.so; do echo "lib=$v"; NEXAR_LIB=$PWD/$v $B 2>&1 | tail -1; done
for nb in 1 3 8; do echo "bands=$nb"; NEXAR_FAST_BANDS=$nb $B 2>&1 | tail -1; done
