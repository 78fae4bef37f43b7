#!/bin/bash
# A/B: custom-mode bench line per library given (NEXAR_LIB), after the parity tests on the default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
for lib in "$@"; do
  n=$(basename $lib .so)
  NEXAR_LIB=$PWD/$lib timeout 120 python bench.py --steps 100 --warmup 10 --mode ${MODE:-custom} --no-cpu-baseline --no-e2e > gpurun_out/ab_${n}.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_issued.avg.per_cycle_active --clock-control none -k regex:"resize|colour|geometry|blur|frame_stats|fixup" -s 12 -c 6 --csv --log-file gpurun_out/launches_custom.csv python bench.py --steps 3 --warmup 3 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
true
