#!/bin/bash
# round-2 ncu captures: strip LU (C5 regime), multi-warp narrow LU + slot solve (C4 regime), banded x banded (C3), launch list
mkdir -p gpurun_out
exec > gpurun_out/prof_r2.log 2>&1
set -x
NCU="ncu --clock-control none --set full"
cap() {  # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 400 $NCU -k regex:$rx -s $skip -c 1 -o gpurun_out/p_$name -f "$@" > /dev/null 2>&1
  ncu -i gpurun_out/p_$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  rm -f gpurun_out/p_$name.ncu-rep
}
cap strip_r2 gbtrf_strip_kernel 1 python tools/prof_case.py widelu 16384 1024 dom
cap mw_r2 gbtrf_mw 1 python tools/prof_case.py lu 131072
cap slotf_r2 'gbtrs_slot<.*false' 1 python tools/prof_case.py lu 131072
cap slotb_r2 'gbtrs_slot<.*true' 1 python tools/prof_case.py lu 131072
cap gbmm_r2 gbmm_bb 1 python tools/prof_case.py gbmm 1048576
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench_r2.log 2>&1
ls -la gpurun_out/*_r2*
