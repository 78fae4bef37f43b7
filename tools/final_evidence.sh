#!/bin/bash
# Round evidence (run under gpurun, one GPU): parity tests, benches, launch list and one full ncu capture per mode.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke=$? >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
python bench.py --mode train --no-cpu-baseline > gpurun_out/bench_train.log 2>&1
python bench.py --mode val --no-cpu-baseline --no-e2e > gpurun_out/bench_val.log 2>&1
python bench.py --mode val --out-dtype f32 --no-cpu-baseline --no-e2e > gpurun_out/bench_val_f32.log 2>&1
python bench.py --workload cfg3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3.log 2>&1
python bench.py --workload cfg4 --steps 30 > gpurun_out/bench_cfg4.log 2>&1
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resize|colour|geometry|blur|frame_stats" -s 15 -c 20 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resize_fast -s 3 -c 1 -o gpurun_out/prof_k1_custom $CMD > gpurun_out/ncu_full.log 2>&1
CMDV="python bench.py --steps 4 --warmup 3 --mode val --no-cpu-baseline --no-e2e"
$CMDV > gpurun_out/plain_val.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resize_fast -s 3 -c 1 -o gpurun_out/prof_k1_val $CMDV > gpurun_out/ncu_full_val.log 2>&1
