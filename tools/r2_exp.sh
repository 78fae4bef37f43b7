#!/bin/bash
# timing experiments: bench lines for each experiment library, then an ncu capture of the fused kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "$@"; do
  n=$(basename $lib .so)
  NEXAR_LIB=$PWD/$lib timeout 120 python bench.py --steps 100 --warmup 10 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/exp_${n}.log 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:resize_fast -s 4 -c 1 -f -o gpurun_out/fused python bench.py --steps 3 --warmup 3 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_issued.avg.per_cycle_active --clock-control none -k regex:"resize|colour|geometry|blur|frame_stats" -s 10 -c 10 --csv --log-file gpurun_out/launches_custom.csv python bench.py --steps 3 --warmup 3 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
true
