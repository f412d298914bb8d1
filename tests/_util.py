import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(name):
    """Yield dicts of the fields saved by tests/golden/make_golden.py for <name>.npz."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    ids = sorted({k.split("_", 1)[0] for k in z.files}, key=lambda s: int(s[1:]))
    for cid in ids:
        yield cid, {k.split("_", 1)[1]: z[k] for k in z.files if k.startswith(cid + "_")}


def scalar(v):
    return v.item() if hasattr(v, "item") else v


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def kat_matrix():
    """The reference's deterministic integer KAT (test/test_linalg.jl:54-56):
    B = BandedMatrix(Symmetric(BandedMatrix{T}(0=>1:10, 1=>11:19, 2=>21:28))), v = 1:10, X = reshape(1:100,10,10)."""
    D = np.zeros((10, 10))
    for i in range(10):
        D[i, i] = 1 + i
    for i in range(9):
        D[i, i + 1] = D[i + 1, i] = 11 + i
    for i in range(8):
        D[i, i + 2] = D[i + 2, i] = 21 + i
    v = np.arange(1.0, 11.0)
    X = np.arange(1.0, 101.0).reshape(10, 10, order="F")
    return D, v, X
