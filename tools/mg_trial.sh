#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/mgt_cfg3_n2.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29520 tools/train_step_share.py --steps 5 > gpurun_out/mgt_cfg5_n2.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/mgt_cfg2_n2.log 2>&1
true
