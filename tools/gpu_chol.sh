#!/bin/bash
# Cholesky pass: parity tests + timings
mkdir -p gpurun_out
exec > gpurun_out/chol.log 2>&1
set -x
timeout 600 python -m pytest tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -5
for kd in 9 10 16 22 31 40 64; do timeout 200 python tools/time_chol.py 524288 $kd U 1 | head -1; done
