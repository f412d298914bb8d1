"""ldiv!(transpose(F), B) timing: python tools/time_gbtrs_t.py n l u nrhs [dom]"""
import sys

import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

n, l, u, nrhs = (int(v) for v in sys.argv[1:5])
A = bm.brand(n, n, l, u, seed=4)
if len(sys.argv) > 5:
    A.data[:, u] += 2.0 * (l + u + 1)
F = bm.lu(A)
for tag, fact in (("N", F), ("T", F.T)):
    ts = []
    for r in range(3):
        X = bm.colmajor(n, nrhs, fill=1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bm.ldiv_(fact, X); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"gbtrs '{tag}' n={n} ({l},{u}) nrhs={nrhs}{' dom' if len(sys.argv) > 5 else ''}: {min(ts[1:]):.2f} ms ({1e6*min(ts[1:])/n/2:.0f} ns/col/sweep)")
