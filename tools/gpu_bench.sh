#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/bench_run.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python bench.py --extras > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
tail -3 gpurun_out/bench_final.err
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
