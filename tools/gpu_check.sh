#!/bin/bash
# standard GPU check (run under gpurun): parity tests, val + custom bench, per-kernel launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 --mode val --no-cpu-baseline --no-e2e > gpurun_out/bench_val.log 2>&1
python bench.py --steps 10 --warmup 3 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/bench_custom.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_issued.avg.per_cycle_active --clock-control none -k regex:"resize|colour|geometry|blur|frame_stats" -s 10 -c 5 --csv --log-file gpurun_out/launches_custom.csv python bench.py --steps 3 --warmup 3 --mode custom --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
