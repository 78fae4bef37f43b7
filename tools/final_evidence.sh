#!/bin/bash
# Round evidence (run under gpurun, one GPU): parity tests, smoke, every bench line, launch lists and full ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo pytest=$? >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke=$? >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 300 python bench.py --mode train --no-cpu-baseline > gpurun_out/bench_train.log 2>&1
timeout 200 python bench.py --mode val --no-cpu-baseline --no-e2e > gpurun_out/bench_val.log 2>&1
timeout 200 python bench.py --mode val --out-dtype f32 --no-cpu-baseline --no-e2e > gpurun_out/bench_val_f32.log 2>&1
timeout 200 python bench.py --out-dtype f32 --no-cpu-baseline --no-e2e > gpurun_out/bench_custom_f32.log 2>&1
timeout 300 python bench.py --workload cfg3 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3.log 2>&1
timeout 300 python bench.py --workload cfg4 --steps 30 > gpurun_out/bench_cfg4.log 2>&1
timeout 300 python tools/train_step_share.py > gpurun_out/cfg5_1gpu.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_issued.avg.per_cycle_active"
K="resize|colour|geometry|blur|frame_stats|fixup|nv12|gather"
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
for m in custom val; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:"$K" -s 12 -c 12 --csv --log-file gpurun_out/launches_$m.csv $CMD --mode $m > gpurun_out/ncu_l_$m.log 2>&1
  timeout 300 ncu --cache-control none --metrics $M --clock-control none -k regex:"$K" -s 12 -c 12 --csv --log-file gpurun_out/insitu_$m.csv $CMD --mode $m > gpurun_out/ncu_i_$m.log 2>&1
done
timeout 300 ncu --metrics $M --clock-control none -k regex:"$K" -s 4 -c 8 --csv --log-file gpurun_out/launches_cfg3.csv $CMD --workload cfg3 --steps 2 > gpurun_out/ncu_l_cfg3.log 2>&1
# full captures: summarised on the box (the .ncu-rep files are too large to travel back), one text file per kernel
full() {  # name, kernel regex, skip, extra bench args
  local n=$1 k=$2 s=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o /tmp/prof_$n $CMD "$@" > gpurun_out/ncu_f_$n.log 2>&1
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > /tmp/raw_$n.csv 2>/dev/null
  ncu -i /tmp/prof_$n.ncu-rep --page source --csv > /tmp/src_$n.csv 2>/dev/null
  { echo "# ncu --set full --clock-control none, kernel regex $k, $CMD $*"; python tools/ncu_summary.py /tmp/raw_$n.csv /tmp/src_$n.csv; } > gpurun_out/summary_$n.txt 2>&1
  rm -f /tmp/prof_$n.ncu-rep /tmp/raw_$n.csv /tmp/src_$n.csv
}
full k1_custom resize_fast 3
full k2_colour colour 3
full k3_geometry geometry 3
full k1_val resize_fast 3 --mode val
full k1_cfg3 resize_fast 1 --workload cfg3 --steps 2
full k3_cfg3 geometry 1 --workload cfg3 --steps 2
true
