#!/bin/bash
# session-3 ncu captures (shipped versions): the three kernels of the blocked Cholesky, the narrow window kernel, launch list
mkdir -p gpurun_out
exec > gpurun_out/prof_s3.log 2>&1
set -x
NCU="ncu --clock-control none --set full"
cap() {  # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 400 $NCU -k regex:$rx -s $skip -c 1 -o gpurun_out/p_$name -f "$@" > /dev/null 2>&1
  ncu -i gpurun_out/p_$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  rm -f gpurun_out/p_$name.ncu-rep
}
cap potf2_r2 pb_potf2_reg 40 python tools/prof_case.py chol 8192 1024
cap trsm_r2 pb_trsm 40 python tools/prof_case.py chol 8192 1024
cap syrk_r2 pb_syrk 40 python tools/prof_case.py chol 8192 1024
cap pbtf2_r2 pbtf2_window 1 python tools/prof_case.py chol 65536 16
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/chol_launches.csv python tools/time_chol.py 4096 1024 U 1 | tail -2
ls -la gpurun_out/*_r2_raw.csv
