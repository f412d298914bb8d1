#!/bin/bash
# 2-GPU sanity: sharded parity tests + the scaling bench line (headline only, then with the other configs)
mkdir -p gpurun_out
exec > gpurun_out/multi.log 2>&1
set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-configs --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','roofline','sharded_check') if k in d})"
