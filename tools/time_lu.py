"""Quick timing of band LU / solve kernels on one GPU (development aid; bench.py --extras is the reported run).
usage: python tools/time_lu.py n l u nrhs [reps]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

n, l, u, nrhs = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
A = bm.brand(n, n, l, u, seed=4)
if len(sys.argv) > 6 and sys.argv[6] == "dom":
    A.data[:, u] += 2.0 * (l + u + 1)  # diagonally dominant: no interchanges
ev = lambda: torch.cuda.Event(enable_timing=True)
tf, ts = [], []
for r in range(reps + 1):
    a, b = ev(), ev()
    a.record()
    F = bm.lu(A)
    b.record()
    b.synchronize()
    tf.append(a.elapsed_time(b))
    X = bm.colmajor(n, nrhs, fill=1.0)
    a.record()
    bm.ldiv_(F, X)
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
R = bm.colmajor(n, nrhs, fill=1.0)
bm.mul_(R, A, X, -1.0, 1.0) if nrhs > 1 else None
res = float(R.abs().max()) if nrhs > 1 else float("nan")
print(f"n={n} (l,u)=({l},{u}) nrhs={nrhs}: lu {min(tf[1:]):.2f} ms ({1e6*min(tf[1:])/n:.1f} ns/col)  "
      f"solve {min(ts[1:]):.2f} ms ({1e6*min(ts[1:])/n/2:.1f} ns/step/sweep)  max|b-Ax|={res:.2e}")
