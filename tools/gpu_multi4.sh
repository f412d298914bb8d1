#!/bin/bash
# 4-GPU check: sharded gbmv parity at world size 2 and 4 + the scaling bench lines
mkdir -p gpurun_out
exec > gpurun_out/multi4.log 2>&1
set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -5
for g in 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $g --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n$g.json
cat gpurun_out/bench_n$g.json
done
