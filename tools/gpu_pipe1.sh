#!/bin/bash
# first GPU pass of the pipelined wide-band LU + the rewritten gbmm kernel
mkdir -p gpurun_out
exec > gpurun_out/pipe1.log 2>&1
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python -m pytest tests/test_gpu_gbmm.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python tools/debug_pipe.py 600 64 64 1 dom
timeout 120 python tools/debug_pipe.py 600 64 64 1
BMB200_PIPE_MAXPANELS=4 timeout 120 python tools/debug_pipe.py 1500 150 140 2
timeout 120 python tools/debug_pipe.py 1500 150 140 2
timeout 120 python tools/debug_pipe.py 2500 300 200 3
timeout 200 python tools/debug_pipe.py 3000 1024 1024 4 dom
timeout 200 python tools/debug_pipe.py 3000 1024 1024 4
timeout 600 python -m pytest tests/test_gpu_lu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/time_lu.py 65536 1024 1024 1 1
BMB200_GBTRF_NOPIPE=1 timeout 300 python tools/time_lu.py 16384 1024 1024 1 1
timeout 300 python tools/prof_case.py gbmm 4194304 && echo gbmm-ok
timeout 600 python bench.py --extras --no-e2e --no-cpu --steps 5 > gpurun_out/bench_pipe1.json 2> gpurun_out/bench_pipe1.err; tail -c 1600 gpurun_out/bench_pipe1.json; tail -5 gpurun_out/bench_pipe1.err
