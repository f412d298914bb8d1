#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/gbmm_wide.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_gbmm.py tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/time_gbmm.py 65536 1024
timeout 200 python tools/time_gbmm.py 262144 256
timeout 200 python tools/time_gbmm.py 1048576 128
timeout 200 python tools/time_chol.py 131072 1024 L 1
