#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/solve4.log 2>&1
timeout 900 ncu --clock-control none --set full --import-source on -k regex:gbtrs_wide_noswap -c 1 -o gpurun_out/solve_full -f python tools/prof_case.py widelu 32768 1024 dom > gpurun_out/ncu_solve.log 2>&1
ncu -i gpurun_out/solve_full.ncu-rep --page raw --csv > gpurun_out/solve_raw.csv 2>/dev/null
ncu -i gpurun_out/solve_full.ncu-rep --page source --csv > gpurun_out/solve_source.csv 2>/dev/null
rm -f gpurun_out/solve_full.ncu-rep
