#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/solve1.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -x -q 2>&1 | tail -12
BMB200_PIPE_STATS=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1 dom
timeout 300 python tools/time_lu.py 65536 1024 1024 4 1 dom
BMB200_GBTRS_NOBLOCK=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1 dom
