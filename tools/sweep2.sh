#!/bin/bash
cd "$(dirname "$0")/.."
B="python bench.py --steps 50 --warmup 5 --mode custom --no-cpu-baseline --no-e2e"
for v in build/*.so; do echo "lib=$v"; NEXAR_LIB=$PWD/$v $B 2>&1 | tail -1; done
