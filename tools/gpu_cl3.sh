#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/cl3.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -x -q -k "wide_band_dominant" 2>&1 | grep -E "Error|assert|rc=|passed|failed" | head -20
timeout 300 python tools/time_lu.py 65536 300 200 8 1 dom
timeout 300 python tools/time_lu.py 2600 300 200 3 1 dom
timeout 300 python tools/time_lu.py 2600 300 200 3 1 dom
