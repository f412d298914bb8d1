#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/variants.log 2>&1
for v in 2_2 2_3 3_2 4_2; do
  cp gpurun_variants/lib_$v.so bandedmatrices.jl_b200/libbmb200.so
  echo "== variant NST_MINB=$v"
  timeout 200 python tools/time_gbmm.py 65536 1024
  timeout 200 python tools/time_gbmm.py 262144 256
done
