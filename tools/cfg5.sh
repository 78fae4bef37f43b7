#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python tools/train_step_share.py --steps 20 > gpurun_out/cfg5_n1.log 2>&1
timeout 300 python tools/train_step_share.py --steps 20 --sync-every-step > gpurun_out/cfg5_n1_sync.log 2>&1
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
timeout 300 $TR --nproc-per-node $N --master-port 29520 tools/train_step_share.py --steps 20 > gpurun_out/cfg5_n$N.log 2>&1
fi
true
