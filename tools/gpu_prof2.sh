#!/bin/bash
# Profiling pass over the secondary kernels (gbmm DMMA, gbtrs/gbtrf register kernels, wide-band LU launch split).
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:gbmm_bb_dmma -s 1 -c 1 -o gpurun_out/gbmm_c3_full -f python tools/prof_case.py gbmm 1048576 > gpurun_out/ncu_gbmm.log 2>&1
ncu -i gpurun_out/gbmm_c3_full.ncu-rep --page raw --csv > gpurun_out/gbmm_c3_raw.csv 2>/dev/null
ncu -i gpurun_out/gbmm_c3_full.ncu-rep --page source --csv > gpurun_out/gbmm_c3_source.csv 2>/dev/null
timeout 600 $NCU --set full --import-source on -k regex:'gbtr[sf]' -s 4 -c 4 -o gpurun_out/lu_c4_full -f python tools/prof_case.py lu 65536 > gpurun_out/ncu_lu.log 2>&1
ncu -i gpurun_out/lu_c4_full.ncu-rep --page raw --csv > gpurun_out/lu_c4_raw.csv 2>/dev/null
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/launches_widelu.csv python tools/prof_case.py widelu 8192 1024 > gpurun_out/ncu_widelu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_widelu.csv', errors='ignore')) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r[4].split('(')[0][:60]
    agg[k][0] += 1; agg[k][1] += float(r[-1])
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={c:6d} total={t/1e6:9.3f} ms avg={t/c/1e3:8.2f} us")
PY
rm -f gpurun_out/launches_widelu.csv
python tools/time_lu.py 65536 1024 1024 1 1
ls -la gpurun_out
