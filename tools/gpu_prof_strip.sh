#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/prof_strip.log 2>&1
set -x
NCU="ncu --clock-control none --set full --import-source on"
timeout 900 $NCU -k regex:gbtrf_strip_kernel -s 1 -c 1 -o gpurun_out/p_strip -f python tools/prof_case.py widelu 16384 1024 dom > /dev/null 2>&1
ncu -i gpurun_out/p_strip.ncu-rep --page raw --csv > gpurun_out/strip_raw.csv 2>/dev/null
ncu -i gpurun_out/p_strip.ncu-rep --page source --csv > gpurun_out/strip_source.csv 2>/dev/null
rm -f gpurun_out/p_strip.ncu-rep
ls -la gpurun_out/strip_*.csv
