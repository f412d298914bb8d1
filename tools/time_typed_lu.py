"""Timing of the typed (S / C / Z) band LU + solve (correctness-first generic kernels): python tools/time_typed_lu.py n kl ku"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

n, kl, ku = (int(v) for v in sys.argv[1:4])
for dt in (torch.float32, torch.complex64, torch.complex128):
    AB = torch.randn((n, 2 * kl + ku + 1), dtype=dt, device="cuda")
    AB[:, :kl] = 0
    AB[:, kl + ku] += 3.0
    W = AB.clone()
    bm.gbtrf_(n, kl, ku, W)
    W = AB.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, piv, info = bm.gbtrf_(n, kl, ku, W)
    tf = time.perf_counter() - t0
    B = torch.randn((4, n), dtype=dt, device="cuda").T
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bm.gbtrs_("N", kl, ku, n, W, piv, B); e1.record(); e1.synchronize()
    print(f"{dt}: n={n} ({kl},{ku}) gbtrf {tf*1e3:.1f} ms ({tf/n*1e9:.0f} ns/col), gbtrs 4 RHS {e0.elapsed_time(e1):.1f} ms ({e0.elapsed_time(e1)/n*1e6/2:.0f} ns/col/sweep), info={info}")
