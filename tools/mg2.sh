#!/bin/bash
# 2-GPU sanity of the torchrun path (cfg2 default line with e2e, cfg3 sharded)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke=$? >> gpurun_out/smoke.log
timeout 300 $TR --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/mg2_cfg2.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/mg2_cfg3.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29503 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/mg2_ref.log 2>&1
true
