"""Small driver used under ncu: runs one configuration of a kernel family once (after a warm-up)."""
import sys

import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm

what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
if what == "lu":
    l, u, nrhs = 16, 16, 64
    A = bm.brand(n, n, l, u, seed=4)
    for _ in range(2):
        F = bm.lu(A)
        X = bm.colmajor(n, nrhs, fill=1.0)
        bm.ldiv_(F, X)
    torch.cuda.synchronize()
elif what == "gbmm":
    A = bm.brand(n, n, 32, 32, seed=2)
    B = bm.brand(n, n, 32, 32, seed=3)
    C = bm.BandedMatrix.undef((n, n), (64, 64))
    for _ in range(2):
        bm.mul_(C, A, B)
    torch.cuda.synchronize()
elif what == "widelu":
    l = u = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    A = bm.brand(n, n, l, u, seed=5)
    if len(sys.argv) > 4 and sys.argv[4] == "dom":
        A.data[:, u] += 2.0 * (l + u + 1)
    for _ in range(2):
        F = bm.lu(A)
        x = torch.ones(n, dtype=torch.float64, device="cuda")
        bm.ldiv_(F, x)
    torch.cuda.synchronize()
