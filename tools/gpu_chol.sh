#!/bin/bash
# Cholesky pass: parity tests + timings
mkdir -p gpurun_out
exec > gpurun_out/chol.log 2>&1
set -x
timeout 600 python -m pytest tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -8
timeout 200 python tools/time_chol.py 131072 1024 U 1 2>&1 | grep -v "panel 10"
timeout 200 python tools/time_chol.py 131072 1024 L 1 2>&1 | grep -v "panel 10"
timeout 200 python tools/time_chol.py 16384 1024 U 1 2>&1 | tail -3
