#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
for i in 1 2; do
for lib in libnexar_clip_b200 exp_q15; do
  NEXAR_LIB=$PWD/vision_collision_detection_b200/$lib.so timeout 120 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e >> gpurun_out/ab_$lib.log 2>&1
done
done
M="gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_issued.avg.per_cycle_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
timeout 300 ncu --metrics $M --clock-control none -k regex:"colour|geometry" -s 4 -c 4 --csv --log-file gpurun_out/launches_q16.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_q16.log 2>&1
true
