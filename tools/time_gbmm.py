"""Quick timing of the banded x banded kernel (development aid). usage: python tools/time_gbmm.py [n] [l] [beta]"""
import sys

import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
l = int(sys.argv[2]) if len(sys.argv) > 2 else 32
beta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
A = bm.brand(n, n, l, l, seed=2)
B = bm.brand(n, n, l, l, seed=3)
C = bm.BandedMatrix.undef((n, n), (2 * l, 2 * l))
C.data.zero_()
ts = []
for r in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    bm.mul_(C, A, B, 1.0, beta)
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b))
ms = min(ts[1:])
W = 2 * l + 1
print(f"gbmm n={n} ({l},{l})x({l},{l}) beta={beta}: {ms:.3f} ms  {2.0*W*W*n/ms/1e9:.2f} TFLOP/s  {8.0*n*(2*W+2*W-1)/ms/1e6:.0f} GB/s algorithmic")
