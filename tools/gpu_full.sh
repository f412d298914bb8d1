#!/bin/bash
# full parity run + extras bench
mkdir -p gpurun_out
exec > gpurun_out/full.log 2>&1
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 900 python bench.py --extras --no-e2e --no-cpu --steps 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1400 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
