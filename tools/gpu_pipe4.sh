#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/pipe4.log 2>&1
set -x
timeout 120 python tools/debug_pipe.py 600 64 64 1 dom
timeout 120 python tools/debug_pipe.py 600 64 64 1
timeout 120 python tools/debug_pipe.py 1500 150 140 2
timeout 120 python tools/debug_pipe.py 2500 300 200 3
timeout 200 python tools/debug_pipe.py 3000 1024 1024 4 dom
timeout 200 python tools/debug_pipe.py 3000 1024 1024 4
timeout 600 python -m pytest tests/test_gpu_lu.py -m gpu -x -q 2>&1 | tail -8
BMB200_PIPE_STATS=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1
BMB200_PIPE_STATS=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1 dom
