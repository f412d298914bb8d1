#!/bin/bash
# Cholesky pass: parity tests + timings (+ the stamped debug build's timeline when tools/_dbg_libbmb200.so is there)
mkdir -p gpurun_out
exec > gpurun_out/chol.log 2>&1
set -x
timeout 600 python -m pytest tests/test_gpu_chol.py -m gpu -x -q 2>&1 | tail -5
timeout 200 python tools/time_chol.py 131072 1024 U 1 | head -3
timeout 200 python tools/time_chol.py 131072 1024 L 1 | head -3
timeout 200 python tools/time_chol.py 131072 256 U 1 | head -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
if [ -f tools/_dbg_libbmb200.so ]; then
  cp bandedmatrices.jl_b200/libbmb200.so /tmp/rel.so
  cp tools/_dbg_libbmb200.so bandedmatrices.jl_b200/libbmb200.so
  timeout 200 python tools/time_chol.py 4096 1024 U 1 | grep STAMP | tail -28
  cp /tmp/rel.so bandedmatrices.jl_b200/libbmb200.so
fi
