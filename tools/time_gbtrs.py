"""A/B timing of the slot-scheduled band solve (csrc/gbtrs_slot.cu) at the C4 shape: every (P, W) variant through the
internal hook, each checked bit-for-bit against OpenBLAS dgbtrs_ on a few right-hand sides.  Development aid.
usage: python tools/time_gbtrs.py [n] [l] [u] [nrhs]   -> one JSON line per variant (also appended to gpurun_out/gbtrs_ab.jsonl)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bandedmatrices_b200 as bm  # noqa: E402
import oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
l = int(sys.argv[2]) if len(sys.argv) > 2 else 16
u = int(sys.argv[3]) if len(sys.argv) > 3 else 16
nrhs = int(sys.argv[4]) if len(sys.argv) > 4 else 256
NCHK = min(4, nrhs)

hd = bm.handle(0)
A = bm.brand(n, n, l, u, seed=4)
F = bm.lu(A)
ab = F.factors.banddata_host()          # (2l+u+1, n) host copy of the factors
ipiv = np.ascontiguousarray(F.ipiv)
rng = np.random.default_rng(3)
Bh = np.asfortranarray(rng.random((n, NCHK)))
ref = Bh.copy(order="F")
ob = oracle.backend("OB")
oracle.ldiv(ob, "N", ab, ipiv, l, u, ref)
Bd = torch.rand((nrhs, n), dtype=torch.float64, device="cuda").T  # column-major n x nrhs
Bd[:, :NCHK] = torch.as_tensor(Bh).cuda()
X = bm.colmajor(n, nrhs)
out = open(os.path.join(ROOT, "gpurun_out", "gbtrs_ab.jsonl"), "a") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None


def run(label, fn):
    best = 1e30
    for r in range(3):
        X.copy_(Bd)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    same = bool(np.array_equal(X[:, :NCHK].cpu().numpy(), ref))
    rec = {"variant": label, "n": n, "l": l, "u": u, "nrhs": nrhs, "ms": round(best, 3),
           "cycles_per_step_pair_at_1965MHz": round(best * 1e-3 * 1.965e9 / n, 1), "bit_identical": same}
    print(json.dumps(rec), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")
        out.flush()


run("dispatch (bmb200_dgbtrs)", lambda: bm.ldiv_(F, X))
dip = F.d_ipiv()
for PF, PB, RF, RB in ((4, 8, 1, 1), (4, 4, 1, 1), (8, 8, 1, 1), (4, 4, 2, 2)):
    if l + PF > 32:
        continue
    for W in (1, 2, 4):
        def f(PF=PF, PB=PB, RF=RF, RB=RB, W=W):
            hd.check(hd.lib.bmb200_internal_gbtrs_slot(hd.h, PF, PB, W, RF, RB, n, l, u, nrhs, C.c_void_p(F.factors.ptr), F.factors.lda,
                                                       C.c_void_p(dip.data_ptr()), C.c_void_p(X.data_ptr()), n), "slot")
        run(f"slot PF={PF} PB={PB} RF={RF} RB={RB} W={W}", f)
