"""Size-matched copy roofline: torch b.copy_(a) moving the same number of bytes (read + write) as gbmv (4,3) at n = 2^27 .. 2^23,
timed the same way (back-to-back launches, CUDA events).  Says how much of the small-n efficiency loss of the strong-scaling runs is
the kernel and how much is the size."""
import torch

for lg in (27, 26, 25, 24, 23):
    nbytes = 80 * (1 << lg)          # bytes gbmv moves at this n
    m = nbytes // 2 // 8             # doubles per buffer: read m, write m
    a = torch.rand(m, dtype=torch.float64, device="cuda")
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    torch.cuda.synchronize()
    K = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        b.copy_(a)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"copy of 80*2^{lg} bytes: {ms*1e3:.1f} us, {nbytes/ms/1e6:.0f} GB/s")
    del a, b
