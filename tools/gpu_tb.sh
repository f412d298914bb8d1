#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/tb.log 2>&1
set -x
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
for kk in 6 8 12 15; do
BMB200_SBMV_ROWS_K=16 timeout 300 python tools/time_tb.py 16777216 $kk 2>&1 | grep "sbmv U"
BMB200_SBMV_ROWS_K=0 timeout 300 python tools/time_tb.py 16777216 $kk 2>&1 | grep "sbmv U"
done
