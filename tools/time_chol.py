"""Timing of the banded Cholesky (pbtrf! / pbtrs!) on the device: python tools/time_chol.py n kd [uplo] [nrhs]."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
kd = int(sys.argv[2]) if len(sys.argv) > 2 else 16
uplo = sys.argv[3] if len(sys.argv) > 3 else "U"
nrhs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
for kv in sys.argv[5:]:  # tuning knobs of the handle, key=value (csrc/common.cuh bmb_tuning)
    key, val = kv.split("=")
    bm.handle(0).tune(key, int(val))
g = torch.Generator(device="cuda").manual_seed(1)
d0 = torch.rand((n, kd + 1), dtype=torch.float64, device="cuda", generator=g) - 0.5
d0[:, kd if uplo == "U" else 0] = 2.0 * (kd + 1)
best = 1e30
for it in range(5):
    d = d0.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, info = bm.pbtrf_(uplo, n, kd, d)  # synchronises (info)
    best = min(best, time.perf_counter() - t0)
    assert info == 0
flops = n * (kd + 1.0) ** 2
print(f"pbtrf {uplo} n={n} kd={kd}: {best * 1e3:.2f} ms, {best / n * 1e9:.1f} ns/column, {flops / best / 1e12:.3f} TFLOP/s")
B = torch.ones((nrhs, n), dtype=torch.float64, device="cuda").T if nrhs > 1 else torch.ones(n, dtype=torch.float64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
bm.pbtrs_(uplo, n, kd, d, B)
ev[0].record()
bm.pbtrs_(uplo, n, kd, d, B)
ev[1].record()
torch.cuda.synchronize()
print(f"pbtrs {uplo} nrhs={nrhs}: {ev[0].elapsed_time(ev[1]):.2f} ms")
