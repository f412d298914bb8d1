#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/mm.log 2>&1
set -x
timeout 900 python -m pytest tests/test_gpu_gbmm.py -m gpu -x -q 2>&1 | tail -3
for rw in 12 16; do
BMB200_GBMM_RW=$rw timeout 120 python tools/time_gbmm.py 4194304 64
BMB200_GBMM_RW=$rw timeout 120 python tools/time_gbmm.py 4194304 48
BMB200_GBMM_RW=$rw BMB200_GBMM_RING=1 timeout 120 python tools/time_gbmm.py 4194304 32
done
