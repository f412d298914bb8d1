#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/cl3.log 2>&1
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_tb.py -m gpu -x -q -k "wide_band_dominant or laplacian or tbsv or tbmv" 2>&1 | tail -3
for args in "2600 300 200 3" "65536 1024 1024 1"; do
  echo "== $args"; timeout 120 python tools/time_lu.py $args 2 dom 2>&1 | tail -1 | cut -c1-200
  echo "== $args (pair off)"; BMB200_GBTRS_PAIR=0 timeout 120 python tools/time_lu.py $args 2 dom 2>&1 | tail -1 | cut -c1-200
done
timeout 300 python tools/time_tb.py 1048576 1024 2>&1 | grep tbsv
