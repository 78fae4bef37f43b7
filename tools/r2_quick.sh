#!/bin/bash
# quick GPU check: parity tests + bench lines (custom / val [/ cfg3]), launch list of one custom step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
for mode in custom val; do
  timeout 200 python bench.py --steps 100 --warmup 10 --mode $mode --no-cpu-baseline --no-e2e > gpurun_out/bench_$mode.log 2>&1
done
if [ -n "$CFG3" ]; then
timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg3 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3.log 2>&1
fi
if [ -n "$NCU" ]; then
timeout 300 ncu --set full --import-source on --clock-control none -k regex:${NCUK:-resize_fast} -s ${NCUS:-4} -c 1 -f -o gpurun_out/prof python bench.py --steps 3 --warmup 3 --mode ${NCU} --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
fi
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_issued.avg.per_cycle_active --clock-control none -k regex:"resize|colour|geometry|blur|frame_stats|fixup" -s 12 -c 8 --csv --log-file gpurun_out/launches_custom.csv python bench.py --steps 3 --warmup 3 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
true
