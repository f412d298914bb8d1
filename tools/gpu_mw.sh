#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/mw.log 2>&1
timeout 60 python tools/time_lu.py 10000 4 3 1 3 || { echo "TIMEOUT/FAIL small"; exit 1; }
timeout 200 python -m pytest tests/test_gpu_lu.py -m gpu -q -x 2>&1 | tail -6
timeout 100 python tools/time_lu.py 1048576 16 16 16 2
timeout 100 python tools/time_lu.py 262144 8 7 4 2
