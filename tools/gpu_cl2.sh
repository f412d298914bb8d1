#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/cl2.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -x -q -k "wide_band_dominant or laplacian" 2>&1 | tail -5
for c in 4 8 16; do
BMB200_GBTRS_CLUSTER=$c timeout 300 python tools/time_lu.py 65536 1024 1024 1 2 dom
done
BMB200_DEBUG=1 BMB200_GBTRS_CLUSTER=8 BMB200_GBTRS_PFDIST=0 timeout 300 python tools/time_lu.py 65536 1024 1024 1 2 dom
BMB200_GBTRS_CLUSTER=8 timeout 300 python tools/time_lu.py 65536 1024 1024 8 2 dom
BMB200_GBTRS_CLUSTER=8 timeout 300 python tools/time_lu.py 65536 256 256 4 2 dom
BMB200_GBTRS_CLUSTER=8 timeout 300 python tools/time_lu.py 65536 1024 0 1 2 dom
BMB200_GBTRS_CLUSTER=8 timeout 300 python tools/time_lu.py 65536 0 1024 1 2 dom
