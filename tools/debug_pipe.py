"""Development aid: wide-band LU (pipelined kernel) vs the C oracle on one shape, with a mismatch report.
usage: python tools/debug_pipe.py n l u [seed] ; env BMB200_PIPE_MAXPANELS / BMB200_GBTRF_NOPIPE are honoured."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bandedmatrices_b200 as bm
import oracle
from oracle import Band, brand, lu

n, l, u = (int(v) for v in sys.argv[1:4])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dom = len(sys.argv) > 5 and sys.argv[5] == "dom"
rng = np.random.default_rng(seed)
A = brand(rng, n, n, l, u)
if dom:
    A.data[u, :] += 2.0 * (l + u + 1)  # diagonally dominant: no interchanges
C = oracle.backend("C")
t0 = time.time()
ab, ipiv, info = lu(C, A)
t1 = time.time()
try:
    F = bm.lu(bm.BandedMatrix.from_banddata(A.data, n, l, u))
except Exception as e:  # noqa
    print(f"n={n} l={l} u={u}: GPU lu raised {type(e).__name__}: {e}")
    sys.exit(1)
got = F.factors.banddata_host()
pm = np.nonzero(F.ipiv != ipiv)[0]
bad = ~((got == ab) | (np.isnan(got) & np.isnan(ab)))
# ignore out-of-matrix corner slots
kv = l + u
rows = np.arange(ab.shape[0])[:, None] - kv + np.arange(n)[None, :]
bad &= (rows >= 0) & (rows < n)
cols = np.nonzero(bad.any(axis=0))[0]
print(f"n={n} l={l} u={u} dom={dom}: oracle {t1-t0:.1f}s, nontrivial pivots {(ipiv != np.arange(1, n+1)).sum()}, "
      f"pivot mismatches {pm.size} (first {pm[:5]}), bad columns {cols.size} (first {cols[:8]}), "
      f"max abs diff {np.nanmax(np.abs(got - ab)) if cols.size else 0.0:.3e}")
if cols.size:
    c = cols[0]
    br = np.nonzero(bad[:, c])[0]
    print(f"  column {c}: bad band rows {br[:10]} .. (matrix rows {br[:10] - kv + c}); got {got[br[:4], c]} ref {ab[br[:4], c]}")
    sys.exit(2)
if pm.size:
    sys.exit(3)
