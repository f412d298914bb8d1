#!/bin/bash
# quick GPU pass: LU/solve parity + extras bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py -m gpu -x -q > gpurun_out/pytest_lu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lu.log
tail -15 gpurun_out/pytest_lu.log
timeout 900 python bench.py --extras --no-e2e --no-cpu --steps 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
