#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/tb.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_tb.py -m gpu -x -q -k sbmv 2>&1 | tail -3
timeout 300 python tools/time_tb.py 67108864 4 2>&1 | grep sbmv
timeout 300 python tools/time_tb.py 16777216 15 2>&1 | grep sbmv
