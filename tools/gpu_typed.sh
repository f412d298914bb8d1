#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/typed.log 2>&1
timeout 900 python -m pytest tests/test_gpu_typed.py tests/test_gpu_tb.py tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -25
