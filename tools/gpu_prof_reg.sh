#!/bin/bash
# ncu capture (full set + source) of the narrow-band LU kernel and the slot solve at a C4-shaped case
mkdir -p gpurun_out
exec > gpurun_out/prof_reg.log 2>&1
set -x
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:gbtrf_mw -s 1 -c 1 -o gpurun_out/p_reg -f python tools/prof_case.py lu 131072 > /dev/null 2>&1
ncu -i gpurun_out/p_reg.ncu-rep --page raw --csv > gpurun_out/reg_raw.csv 2>/dev/null
ncu -i gpurun_out/p_reg.ncu-rep --page source --csv > gpurun_out/reg_source.csv 2>/dev/null
rm -f gpurun_out/p_reg.ncu-rep
ls -la gpurun_out/reg_*.csv
