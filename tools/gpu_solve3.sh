#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/solve3.log 2>&1
timeout 600 python -m pytest tests/test_gpu_lu.py -m gpu -x -q -k "division or dominant or laplacian or golden" 2>&1 | tail -6
timeout 300 python tools/time_lu.py 65536 1024 1024 1 1 dom 2>&1 | grep -v "tid 992" | tail -3
