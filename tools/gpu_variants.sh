#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/variants.log 2>&1
for v in old new; do
  cp gpurun_variants/lib_$v.so bandedmatrices.jl_b200/libbmb200.so
  echo "== variant $v"
  for k in 16 40 100 300 1024; do python tools/time_tb.py 1048576 $k 2>&1 | grep sbmv; done
done
