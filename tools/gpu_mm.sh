#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/mm.log 2>&1
set -x
for nt in 2 3 4 6; do BMB200_GBMM_NT=$nt timeout 120 python tools/time_gbmm.py; done
