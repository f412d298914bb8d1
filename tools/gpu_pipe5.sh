#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/pipe5.log 2>&1
BMB200_PIPE_STATS=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1 dom
