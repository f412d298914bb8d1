"""gbmv (4,3) at the per-GPU sizes of the 1/2/4/8-GPU strong-scaling runs: back-to-back launches, CUDA events."""
import sys

import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

spr = int(sys.argv[1]) if len(sys.argv) > 1 else 0
bm.handle(0).tune("gbmv_spr", spr)
print("sets per run override:", spr)
for lg in (27, 25, 24, 23):
    n = 1 << lg
    A = bm.brand(n, n, 4, 3, seed=1)
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(5):
        bm.mul_(y, A, x)
    torch.cuda.synchronize()
    K = 40
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        bm.mul_(y, A, x)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / K
    print(f"n=2^{lg}: {ms*1e3:.1f} us/step, {80.0*n/ms/1e6:.0f} GB/s, frac of 6553: {80.0*n/ms/1e6/6553.3:.3f}, excess over ideal {ms*1e3 - 80.0*n/6553.3e3:.1f} us")
    del A, x, y
