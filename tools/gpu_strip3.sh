#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/strip3.log 2>&1
timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
hd = bm.handle(0)
hd.tune("pipe_stats", 1)
n = 4096
A = bm.brand(n, n, 300, 200, seed=5)
A.data[:, 200] += 2.0 * 501
try:
    F = bm.lu(A)
except Exception as e:
    print("ERR", e)
torch.cuda.synchronize()
PY
