#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/pipe3.log 2>&1
set -x
timeout 900 ncu --clock-control none --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on -k regex:gbtrf_pipe_kernel -s 1 -c 1 -o gpurun_out/pipe_src -f python tools/prof_case.py widelu 16384 1024 dom > gpurun_out/ncu_pipe.log 2>&1
tail -5 gpurun_out/ncu_pipe.log
ncu -i gpurun_out/pipe_src.ncu-rep --page source --csv > gpurun_out/pipe_source.csv 2>/dev/null
ncu -i gpurun_out/pipe_src.ncu-rep --page raw --csv > gpurun_out/pipe_raw.csv 2>/dev/null
rm -f gpurun_out/pipe_src.ncu-rep
ls -la gpurun_out
BMB200_PIPE_STATS=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1 dom
