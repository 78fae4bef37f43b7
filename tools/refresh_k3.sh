#!/bin/bash
# refresh the custom-mode launch lists and the geometry kernel's full summary (after a K3-only change)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_issued.avg.per_cycle_active"
K="resize|colour|geometry|blur|frame_stats|fixup|nv12|gather"
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 ncu --metrics $M --clock-control none -k regex:"$K" -s 12 -c 12 --csv --log-file gpurun_out/launches_custom.csv $CMD > gpurun_out/ncu_l_custom.log 2>&1
timeout 300 ncu --cache-control none --metrics $M --clock-control none -k regex:"$K" -s 12 -c 12 --csv --log-file gpurun_out/insitu_custom.csv $CMD > gpurun_out/ncu_i_custom.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:geometry -s 3 -c 1 -f -o /tmp/prof_k3 $CMD > gpurun_out/ncu_f_k3.log 2>&1
ncu -i /tmp/prof_k3.ncu-rep --page raw --csv > /tmp/raw_k3.csv 2>/dev/null
ncu -i /tmp/prof_k3.ncu-rep --page source --csv > /tmp/src_k3.csv 2>/dev/null
{ echo "# ncu --set full --clock-control none, kernel regex geometry, $CMD"; python tools/ncu_summary.py /tmp/raw_k3.csv /tmp/src_k3.csv; } > gpurun_out/summary_k3_geometry.txt 2>&1
true
