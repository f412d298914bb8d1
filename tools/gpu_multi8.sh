#!/bin/bash
# 8-GPU check: the scaling bench line the driver will run at N=8
mkdir -p gpurun_out
exec > gpurun_out/multi8.log 2>&1
set -x
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n8.json
cat gpurun_out/bench_n8.json
