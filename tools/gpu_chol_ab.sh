#!/bin/bash
# A/B of one tuning knob of the blocked Cholesky inside one box: bash tools/gpu_chol_ab.sh key
mkdir -p gpurun_out
exec > gpurun_out/chol_ab.log 2>&1
K=${1:-pb_clate}
for rep in 1 2; do
  for v in 0 1; do
    for kd in 1024 256; do
      echo -n "$K=$v "; timeout 200 python tools/time_chol.py 131072 $kd U 1 $K=$v | head -1
    done
  done
done
