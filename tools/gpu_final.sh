#!/bin/bash
# Round-end GPU pass: parity, smoke, both bench arms, extras, launch list + full ncu captures (summaries go to profiles/).
mkdir -p gpurun_out
exec > gpurun_out/final.log 2>&1
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --extras --no-e2e --steps 5 > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; tail -c 1300 gpurun_out/bench_extras.json
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_gbmv_c2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:gbmv_n_systolic -s 3 -c 1 -o gpurun_out/gbmv_c2_full -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gbmv_c2_full.ncu-rep --page raw --csv > gpurun_out/gbmv_c2_raw.csv 2>/dev/null
timeout 600 $NCU --set full --import-source on -k regex:gbmm_bb_dmma -s 1 -c 1 -o gpurun_out/gbmm_c3_full -f python tools/prof_case.py gbmm 1048576 > gpurun_out/ncu_gbmm.log 2>&1
ncu -i gpurun_out/gbmm_c3_full.ncu-rep --page raw --csv > gpurun_out/gbmm_c3_raw.csv 2>/dev/null
timeout 900 $NCU --set full -k regex:gbtrf_pipe_kernel -s 1 -c 1 -o gpurun_out/pipe_full -f python tools/prof_case.py widelu 16384 1024 dom > gpurun_out/ncu_pipe.log 2>&1
ncu -i gpurun_out/pipe_full.ncu-rep --page raw --csv > gpurun_out/pipe_raw.csv 2>/dev/null
timeout 900 $NCU --set full -k regex:gbtrs_cluster -s 1 -c 1 -o gpurun_out/cluster_full -f python tools/prof_case.py widelu 16384 1024 dom > gpurun_out/ncu_cluster.log 2>&1
ncu -i gpurun_out/cluster_full.ncu-rep --page raw --csv > gpurun_out/cluster_raw.csv 2>/dev/null
timeout 900 $NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/launches_c5.csv python tools/prof_case.py widelu 65536 1024 dom > gpurun_out/ncu_c5.log 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out
