#!/bin/bash
# ncu captures of the HBM-bound kernels added in session 5 (traffic vs algorithmic bytes)
mkdir -p gpurun_out
exec > gpurun_out/prof_new.log 2>&1
set -x
NCU="ncu --clock-control none --set full"
timeout 600 $NCU -k regex:tbmv_sweep -s 1 -c 1 -o gpurun_out/p_tbmv -f python tools/prof_case.py tb 1048576 1024 > /dev/null 2>&1
timeout 600 $NCU -k regex:sbmv_rows -s 1 -c 1 -o gpurun_out/p_sbmv -f python tools/prof_case.py sbmv 134217728 3 > /dev/null 2>&1
timeout 600 $NCU -k regex:band_ewise -s 1 -c 1 -o gpurun_out/p_axpy -f python tools/prof_case.py axpy 67108864 > /dev/null 2>&1
for f in tbmv sbmv axpy; do ncu -i gpurun_out/p_$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/*_raw.csv
