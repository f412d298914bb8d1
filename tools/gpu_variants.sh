#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/variants.log 2>&1
for v in nb64 nb128; do
  cp gpurun_variants/lib_$v.so bandedmatrices.jl_b200/libbmb200.so
  echo "== variant $v"
  timeout 600 python -m pytest tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -3
  timeout 200 python tools/time_chol.py 131072 1024 U 1
  timeout 200 python tools/time_chol.py 131072 1024 L 1
  timeout 200 python tools/time_chol.py 131072 200 U 1
done
