#!/bin/bash
# final pass of the session: every GPU test, smoke, default bench line, extras, launch list of pbtrf + wide product, ncu of the final K-blocked kernel
mkdir -p gpurun_out
exec > gpurun_out/final.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 1200 python bench.py --extras > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
tail -3 gpurun_out/bench_final.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/chol_launches.csv python tools/time_chol.py 4096 1024 U 1 | tail -2
NCU="ncu --clock-control none --set full"
timeout 400 $NCU -k regex:gbmm_bb_kblock -s 1 -c 1 -o gpurun_out/p_kblock -f python tools/prof_case.py widegbmm 16384 1024 > /dev/null 2>&1
ncu -i gpurun_out/p_kblock.ncu-rep --page raw --csv > gpurun_out/kblock_final_raw.csv 2>/dev/null; rm -f gpurun_out/p_kblock.ncu-rep
