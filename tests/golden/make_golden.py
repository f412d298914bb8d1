"""Generates tests/golden/*.npz with the OpenBLAS 0.3.30 ILP64 library bundled in numpy -- the same
Fortran entry points BandedMatrices.jl ccalls (dgbmv_: src/blas.jl:19-26, dgbtrf_: src/banded/BandedLU.jl:98,
dgbtrs_: src/banded/linalg.jl:28), driven with the reference's argument conventions through oracle/.
Julia itself is not installed in this image, so these are "outputs of the reference's arithmetic backend",
not of the Julia package.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import Band, backend, banded_muladd_vec, brand, gbmm_kernel, gbmv, ldiv, lu  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
OB = backend("OB")
rng = np.random.default_rng(12345)

# ---- gbmv: shapes of test/test_broadcasting.jl:353-394 and test/test_banded.jl:97-140 ------------------
cases = {}
k = 0
for (m, n, l, u) in [(100, 100, 1, 1), (100, 100, 0, 1), (100, 100, 1, 0), (100, 100, 0, 0), (10, 12, 2, 3),
                     (12, 10, 3, 2), (257, 257, 4, 3), (300, 300, 32, 32), (64, 64, 9, 5)]:
    for (al, be) in [(1.0, 0.0), (2.0, 3.0)]:
        for tr in "NT":
            A = brand(rng, m, n, l, u, corners=np.nan)
            x = rng.standard_normal(n if tr == "N" else m)
            y0 = rng.standard_normal(m if tr == "N" else n)
            y = y0.copy()
            gbmv(OB, tr, m, l, u, al, A.data, x, be, y)
            cases[f"c{k}"] = dict(trans=tr, m=m, n=n, l=l, u=u, alpha=al, beta=be, data=A.data, x=x, y0=y0, y=y)
            k += 1
np.savez_compressed(os.path.join(OUT, "gbmv.npz"), **{f"{c}_{f}": v for c, d in cases.items() for f, v in d.items()})

# ---- gbmm: a slice of test/test_linalg.jl:212-222 (n,nu,m in {1,5,50}; bands in {0,1,2,30}; 0.123/0.456) ----
cases = {}
k = 0
import itertools

for n, nu, m in [(5, 5, 5), (50, 50, 50), (50, 5, 50), (5, 50, 1), (1, 5, 50), (50, 1, 5)]:
    for Al, Au, Bl, Bu in [(0, 0, 0, 0), (1, 2, 2, 1), (30, 1, 0, 2), (2, 30, 30, 0), (30, 30, 30, 30), (0, 2, 1, 0)]:
        A = brand(rng, n, nu, Al, Au, corners=np.nan)
        B = brand(rng, nu, m, Bl, Bu, corners=np.nan)
        Cl, Cu = min(n - 1, Al + Bl), min(m - 1, Au + Bu)
        C0 = brand(rng, n, m, Cl, Cu)
        c = C0.data.copy(order="F")
        gbmm_kernel(OB, 0.123, A.data, B.data, 0.456, c, n, nu, m, Al, Au, Bl, Bu, Cl, Cu)
        cases[f"c{k}"] = dict(dims=np.array([n, nu, m, Al, Au, Bl, Bu, Cl, Cu]), A=A.data, B=B.data, C0=C0.data, C=c)
        k += 1
np.savez_compressed(os.path.join(OUT, "gbmm.npz"), **{f"{c}_{f}": v for c, d in cases.items() for f, v in d.items()})

# ---- lu + solve: test/test_bandedlu.jl shapes + README / C4-style bands -----------------------------------
cases = {}
k = 0
for (n, l, u, nrhs) in [(5, 1, 1, 1), (10, 2, 1, 10), (100, 4, 3, 1), (200, 16, 16, 3), (150, 5, 7, 5), (90, 0, 3, 2),
                        (90, 3, 0, 2), (130, 64, 64, 2), (400, 33, 31, 4)]:
    A = brand(rng, n, n, l, u)
    ab, ipiv, info = lu(OB, A)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    X = B.copy(order="F")
    ldiv(OB, "N", ab, ipiv, l, u, X)
    XT = B.copy(order="F")
    ldiv(OB, "T", ab, ipiv, l, u, XT)
    cases[f"c{k}"] = dict(dims=np.array([n, l, u, nrhs, info]), data=A.data, ab=ab, ipiv=ipiv, B=B, X=X, XT=XT)
    k += 1
np.savez_compressed(os.path.join(OUT, "lu.npz"), **{f"{c}_{f}": v for c, d in cases.items() for f, v in d.items()})
print("golden fixtures written with", OB.config)
