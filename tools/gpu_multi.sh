#!/bin/bash
# 2-GPU sanity: sharded gbmv parity test + the scaling bench line
mkdir -p gpurun_out
exec > gpurun_out/multi.log 2>&1
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -3
