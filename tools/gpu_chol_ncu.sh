#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/chol_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:pb_ -c 120 --csv --log-file gpurun_out/chol_launches.csv python tools/time_chol.py 16384 1024 U 1
tail -2 gpurun_out/chol_launches.csv
