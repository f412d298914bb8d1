#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/strip2.log 2>&1
timeout 240 python -m pytest tests/test_gpu_lu.py -m gpu -x -q -k "wide or laplacian" 2>&1 | tail -5
timeout 150 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
hd = bm.handle(0)
hd.tune("pipe_stats", 1)
for n in (1 << 16,):
    A = bm.brand(n, n, 1024, 1024, seed=5)
    A.data[:, 1024] += 2.0 * 2049
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); F = bm.lu(A); b.record(); b.synchronize()
        print("n", n, "lu ms", a.elapsed_time(b), flush=True)
    del F
PY
