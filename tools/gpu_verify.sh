#!/bin/bash
# last check of the shipped state: full GPU parity suite + smoke + a short default bench
mkdir -p gpurun_out
exec > gpurun_out/verify.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-900
