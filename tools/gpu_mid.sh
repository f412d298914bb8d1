#!/bin/bash
# mid-session check: full GPU parity suite + extras bench (C1/C3/C4/C5)
mkdir -p gpurun_out
exec > gpurun_out/mid.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --extras --no-e2e --steps 5 > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; tail -c 1500 gpurun_out/bench_extras.json; tail -3 gpurun_out/bench_extras.err
