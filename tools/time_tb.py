"""Timing of the triangular band kernels (development aid). usage: python tools/time_tb.py [n] [k]"""
import sys

import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = torch.rand((n, k + 1), dtype=torch.float64, device="cuda") / (2 * k)
b = torch.ones(n, dtype=torch.float64, device="cuda")
for uplo in "UL":
    d[:, :] = torch.rand((n, k + 1), dtype=torch.float64, device="cuda") / (2 * k)
    d[:, k if uplo == "U" else 0] = 2.0
    for name, fn in (("tbsv", bm.tbsv_), ("tbmv", bm.tbmv_)):
        ts = []
        for r in range(4):
            x = b.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(uplo, "N", "N", n, k, d, x)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts[1:])
        print(f"{name} {uplo} n={n} k={k}: {ms:.3f} ms  {8.0*n*(k+1)/ms/1e6:.0f} GB/s")
for uplo in "UL":
    xx = torch.rand(n, dtype=torch.float64, device="cuda")
    yy = torch.zeros(n, dtype=torch.float64, device="cuda")
    ts = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bm.sbmv_(uplo, k, 1.0, d, xx, 0.0, yy)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(f"sbmv {uplo} n={n} k={k}: {ms:.3f} ms  {8.0*n*(k+3)/ms/1e6:.0f} GB/s algorithmic (stored triangle + x + y once)")
