#!/bin/bash
# full ncu capture of kernel regex $KREG (default geometry) for each experiment library given
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "$@"; do
  n=$(basename $lib .so)
  CMD="python bench.py --steps 4 --warmup 3 --mode ${MODE:-custom} $EXTRA --no-cpu-baseline --no-e2e"
  NEXAR_LIB=$PWD/$lib timeout 300 ncu --set full --clock-control none --import-source on -k regex:${KREG:-geometry} -s ${SKIP:-2} -c 1 -f -o gpurun_out/prof_$n $CMD > gpurun_out/ncu_$n.log 2>&1
done
true
