#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/bench_only.log 2>&1
timeout 1500 python bench.py --extras > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
tail -3 gpurun_out/bench_final.err
