#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/cl3.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_tb.py -m gpu -x -q -k "wide_band_dominant or laplacian or tb" 2>&1 | tail -3
BMB200_GBTRS_STATS=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1 dom 2>&1 | tail -4
timeout 300 python tools/time_lu.py 65536 1024 1024 1 3 dom
timeout 300 python tools/time_tb.py 1048576 1024
