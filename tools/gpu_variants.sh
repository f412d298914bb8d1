#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/variants.log 2>&1
cat > /tmp/w.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
hd = bm.handle(0)
for force in (0, 1):
    hd.tune("gbmm_wide", 1 if force else -1)
    print("forced wide" if force else "default dispatch")
    for n, l in ((1 << 20, 64), (1 << 21, 48), (1 << 22, 32), (1 << 20, 80), (1<<22, 16)):
        sys.argv = ["x", str(n), str(l)]
        exec(open("tools/time_gbmm.py").read())
        print("   path", hd.last_gbmm_path())
PY
python /tmp/w.py
