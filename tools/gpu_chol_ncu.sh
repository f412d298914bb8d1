#!/bin/bash
mkdir -p gpurun_out
exec >> gpurun_out/chol.log 2>&1
set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/chol_launches.csv python tools/time_chol.py 4096 1024 U 1 | tail -3
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/chol_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:40]].append(float(r[vi].replace(",", "")))
for k, v in d.items():
    print(k, len(v), "mean us", sum(v) / len(v) / 1e3, "max", max(v) / 1e3)
PY
