#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/pipe2.log 2>&1
set -x
BMB200_PIPE_STATS=1 timeout 300 python tools/time_lu.py 65536 1024 1024 1 1
BMB200_PIPE_STATS=1 timeout 300 python tools/time_lu.py 32768 256 256 1 1
timeout 600 ncu --clock-control none --set full --import-source on -k regex:gbmm_bb_dmma -s 1 -c 1 -o gpurun_out/gbmm_c3_v3 -f python tools/prof_case.py gbmm 1048576 > gpurun_out/ncu_gbmm.log 2>&1
ncu -i gpurun_out/gbmm_c3_v3.ncu-rep --page raw --csv > gpurun_out/gbmm_c3_v3_raw.csv 2>/dev/null
ncu -i gpurun_out/gbmm_c3_v3.ncu-rep --page source --csv > gpurun_out/gbmm_c3_v3_source.csv 2>/dev/null
rm -f gpurun_out/gbmm_c3_v3.ncu-rep
