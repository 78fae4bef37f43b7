#!/bin/bash
# per-kernel durations with the caches left as the previous kernel left them (one pass, no flush, no clock control)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in ${MODES:-custom}; do
timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"resize|colour|geometry|blur|frame_stats|fixup|nv12" -s ${SKIP:-40} -c ${CNT:-12} --csv --log-file gpurun_out/insitu_$m.csv python bench.py --steps 10 --warmup 10 --mode $m $EXTRA --no-cpu-baseline --no-e2e > gpurun_out/insitu_$m.log 2>&1
done
true
