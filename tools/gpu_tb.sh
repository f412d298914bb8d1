#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/tb.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_ewise.py -m gpu -x -q 2>&1 | tail -15
