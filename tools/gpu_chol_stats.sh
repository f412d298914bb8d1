#!/bin/bash
mkdir -p gpurun_out
exec >> gpurun_out/chol.log 2>&1
cat > /tmp/st.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
from bandedmatrices_b200 import handle
hd = handle(0); hd.tune("pipe_stats", 1)
for n, kd in ((1 << 14, 1024),):
    d = torch.rand((n, kd + 1), dtype=torch.float64, device="cuda") - 0.5
    d[:, kd] = 2.0 * (kd + 1)
    print("n", n, "kd", kd, flush=True)
    bm.pbtrf_("U", n, kd, d)
PY
python /tmp/st.py
