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
elif what == "tb":
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    d = torch.rand((n, k + 1), dtype=torch.float64, device="cuda") / (2 * k)
    d[:, k] = 2.0
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        bm.tbmv_("U", "N", "N", n, k, d, x)
    torch.cuda.synchronize()
elif what == "sbmv":
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    d = torch.rand((n, k + 1), dtype=torch.float64, device="cuda")
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        bm.sbmv_("U", k, 1.0, d, x, 0.0, y)
    torch.cuda.synchronize()
elif what == "axpy":
    X, Y = bm.brand(n, n, 4, 3, seed=1), bm.brand(n, n, 4, 3, seed=2)
    for _ in range(2):
        bm.axpy_(0.5, X, Y)
    torch.cuda.synchronize()
elif what == "widegbmm":
    l = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    A = bm.brand(n, n, l, l, seed=2)
    B = bm.brand(n, n, l, l, seed=3)
    C = bm.BandedMatrix.undef((n, n), (2 * l, 2 * l))
    for _ in range(2):
        bm.mul_(C, A, B)
    torch.cuda.synchronize()
elif what == "chol":
    kd = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    d = torch.rand((n, kd + 1), dtype=torch.float64, device="cuda") - 0.5
    d[:, kd] = 2.0 * (kd + 1)
    for _ in range(2):
        e = d.clone()
        bm.pbtrf_("U", n, kd, e)
    torch.cuda.synchronize()
