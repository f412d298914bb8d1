#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/c4.log 2>&1
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/time_lu.py 1048576 16 16 256 1
timeout 300 python tools/time_lu.py 262144 100 100 8 1
timeout 300 python tools/time_lu.py 65536 1024 1024 1 1
