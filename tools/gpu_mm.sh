#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/mm.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_gbmm.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/time_gbmm.py
timeout 120 python tools/time_gbmm.py 4194304 32 0.5
