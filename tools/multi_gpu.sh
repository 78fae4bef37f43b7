#!/bin/bash
# 8-GPU box: BASELINE configs[2] (one 256-clip batch sharded over 2/4/8 GPUs), configs[4] (DDP step share), and the
# default line at 8 GPUs with both e2e legs (NUMA-bound and not).  Every line lands in gpurun_out/mg_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/mg_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/mg_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/mg_cfg3_n$n.log 2>&1
done
timeout 300 $TR --nproc-per-node 8 --master-port 29520 tools/train_step_share.py --steps 10 > gpurun_out/mg_cfg5_n8.log 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/mg_cfg2_n8.log 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/mg_cfg2_n4.log 2>&1
timeout 300 $TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/mg_cfg2_n2.log 2>&1
true
