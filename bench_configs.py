"""The other BASELINE.json configurations, timed AND verified inside the default bench.py run (reported under "configs"):

  C1  README workload brand(10000,10000,4,3): A*b, A*A, A\\b              (latency; bit-compared with OpenBLAS in full)
  C3  banded x banded n=2^22, (32,32)x(32,32) -> (64,64)                  (HBM roofline; column-sharded at N > 1)
  C4  banded LU + solve n=2^20, (16,16), 256 RHS                          (gbtrf: ns/column; gbtrs: HBM roofline; RHS-sharded at N > 1)
  C4_chol  the C4 shape through the banded Cholesky: pbtrf! + pbtrs!, 256 RHS (factor bit-compared with OpenBLAS in full)
  C3_wide  banded x banded with C5-sized bands, n=2^16, (1024,1024)^2   (FP64 tensor roofline; K-blocked DMMA kernel)
  C5  2-D Laplacian N=1024: n=2^20, l=u=1024, lu + ldiv!                  (FP64 tensor roofline; one GPU)
      + "cholesky": the same SPD system through pbtrf! / pbtrs! (cholesky(Symmetric(A)) and its ldiv!)

Every entry carries: the GPU time (CUDA events on the launching stream, best of a few repetitions after a warm-up), a
`roofline` object (SURVEY.md 8d algorithmic bytes / flops over the measured peak), a `cpu_baseline` (the reference's own call
sequence replayed from C against OpenBLAS 0.3.30 on the box's host cores -- oracle/blasdriver.c -- on a bounded sample with the
extrapolation stated) and a `parity` object computed at the FULL benchmark size against that same OpenBLAS (the oracle is the
checker here, never the thing measured)."""
import ctypes as C
import json
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
i64, dbl, vp = C.c_int64, C.c_double, C.c_void_p


def fp64_peak():
    p = os.path.join(ROOT, "profiles", "fp64_peaks_r1.json")
    if os.path.exists(p):
        return float(json.load(open(p))["dmma_8x8x4_tflops"]), "measured on this pool's B200 (profiles/fp64_peaks_r1.json, DMMA.8x8x4 register-resident loop)"
    return 37.0, "nominal"


def cpu_protos(L):
    L.drv_gbmv.restype = dbl
    L.drv_gbmv.argtypes = [i64] * 4 + [dbl, vp, i64, vp, dbl, vp]
    L.drv_gbmm.restype = dbl
    L.drv_gbmm.argtypes = [i64] * 9 + [dbl, vp, i64, vp, i64, dbl, vp, i64]
    L.drv_gbtrf.restype = dbl
    L.drv_gbtrf.argtypes = [i64] * 4 + [vp, i64, vp, vp]
    L.drv_gbtrs.restype = dbl
    L.drv_gbtrs.argtypes = [i64] * 4 + [vp, i64, vp, vp, i64, vp]
    return L


def _best(fn, reps=3, setup=None, warm=1):
    ts = []
    for i in range(reps + warm):
        if setup:
            setup()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        b.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return min(ts), float(np.mean(ts))


def _host_band(t: torch.Tensor) -> np.ndarray:
    """(ncols, rows) device band slab -> (rows, ncols) Fortran-ordered host array (LAPACK band storage)."""
    return np.asfortranarray(t.detach().cpu().numpy().T)


def _bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    return bool(np.array_equal(a, b, equal_nan=True))


def laplacian(bm, N, ncols=None):
    """examples/finitedifference_2d.jl: A = I - dt*Laplacian_2D, dt = 1/(4 N^2): diag 2, +-1 and +-N bands -0.25 (the +-1 band
    is 0 across a grid-line edge).  ``ncols`` < N*N returns the leading principal block (same entries)."""
    n = N * N if ncols is None else ncols
    A = bm.BandedMatrix.zeros((n, n), (N, N))
    d = A.data  # (n, 2N+1): d[j, r] = band row r of column j
    d[:, N] = 2.0
    j = torch.arange(n, device=d.device)
    d[1:, N - 1] = torch.where(j[1:] % N != 0, -0.25, 0.0).to(d.dtype)
    d[:-1, N + 1] = torch.where((j[:-1] + 1) % N != 0, -0.25, 0.0).to(d.dtype)
    d[N:, 0] = -0.25
    d[:-N, 2 * N] = -0.25
    return A


# ---------------------------------------------------------------------------------------------------------------------
def run_c1(bm, L, hbm_peak):
    n, l, u = 10000, 4, 3
    A = bm.brand(n, n, l, u, seed=1)
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(b)
    Cm = bm.BandedMatrix.undef((n, n), (2 * l, 2 * u))
    t_ab, _ = _best(lambda: bm.mul_(y, A, b), reps=5)
    t_aa, _ = _best(lambda: bm.mul_(Cm, A, A), reps=5)
    keep = {}

    def solve():
        keep["x"] = bm.solve(A, b)

    t_s, _ = _best(solve, reps=5)
    out = {"workload": "README brand(10000,10000,4,3): A*b, A*A, A\\b (lu + ldiv!)", "Ab_us": round(1e3 * t_ab, 1),
           "AA_us": round(1e3 * t_aa, 1), "solve_us": round(1e3 * t_s, 1), "roofline": None,
           "note": "latency-bound (0.8 MB): times and parity only, no roofline claim (SURVEY.md 8d)"}
    if L is not None:
        a_h, b_h = _host_band(A.data), b.cpu().numpy()
        y_h = np.zeros(n)
        L.drv_set_threads(1)
        c_ab = min(L.drv_gbmv(n, n, l, u, 1.0, a_h.ctypes.data, l + u + 1, b_h.ctypes.data, 0.0, y_h.ctypes.data) for _ in range(5))
        c_h = np.zeros((2 * l + 2 * u + 1, n), order="F")
        c_aa = min(L.drv_gbmm(n, n, n, l, u, l, u, 2 * l, 2 * u, 1.0, a_h.ctypes.data, l + u + 1, a_h.ctypes.data, l + u + 1, 0.0,
                              c_h.ctypes.data, 2 * l + 2 * u + 1) for _ in range(3))
        ab = np.zeros((2 * l + u + 1, n), order="F")
        ab[l:, :] = a_h
        ipiv = np.zeros(n, dtype=np.int64)
        info = np.zeros(1, dtype=np.int64)
        x_h = b_h.copy()
        c_f = L.drv_gbtrf(n, n, l, u, ab.ctypes.data, 2 * l + u + 1, ipiv.ctypes.data, info.ctypes.data)
        c_s = L.drv_gbtrs(n, l, u, 1, ab.ctypes.data, 2 * l + u + 1, ipiv.ctypes.data, x_h.ctypes.data, n, info.ctypes.data)
        got_c = _host_band(Cm.data)
        # corner slots of C's band storage are never written by either side: compare the in-matrix entries only
        rr, jj = np.meshgrid(np.arange(c_h.shape[0]), np.arange(n), indexing="ij")
        inm = (jj + rr - 2 * u >= 0) & (jj + rr - 2 * u < n)
        out["parity"] = {"Ab_bit_identical": _bits_equal(y.cpu().numpy(), y_h),
                         "AA_bit_identical": _bits_equal(got_c[inm], c_h[inm]),
                         "solve_bit_identical": _bits_equal(keep["x"].cpu().numpy(), x_h),
                         "against": "OpenBLAS 0.3.30 dgbmv_ / per-column dgbmv_ replay of _gbmm! / dgbtrf_+dgbtrs_, full size"}
        out["cpu_baseline"] = {"Ab_us": round(1e6 * c_ab, 1), "AA_us": round(1e6 * c_aa, 1), "solve_us": round(1e6 * (c_f + c_s), 1),
                               "cores": 1, "kind": "reference", "sample": "full size (OpenBLAS runs these sizes on one thread)"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def run_c3(bm, L, hbm_peak, rank=0, world=1, n=1 << 22):
    import torch.distributed as dist

    Al = Au = Bl = Bu = 32
    A = bm.brand(n, n, Al, Au, seed=2)
    B = bm.brand(n, n, Bl, Bu, seed=3)
    Cm = bm.BandedMatrix.undef((n, n), (Al + Bl, Au + Bu))
    flops = 2.0 * (Al + Au + 1) * (Bl + Bu + 1) * n
    byts = 8.0 * n * ((Al + Au + 1) + (Bl + Bu + 1) + (Al + Bl + Au + Bu + 1))
    tpeak, tsrc = fp64_peak()
    out = {"workload": f"BandedMatrix*BandedMatrix Float64 n=2^{int(np.log2(n))}, (32,32)x(32,32)->(64,64), alpha=1, beta=0",
           "algorithmic_bytes": byts, "flops": flops}
    if world == 1:
        t, tm = _best(lambda: bm.mul_(Cm, A, B), reps=5)
        out.update({"ms": round(t, 3), "ms_mean": round(tm, 3), "GFLOPs": round(flops / t / 1e6, 1), "GBs": round(byts / t / 1e6, 1),
                    "roofline": {"bound": "hbm", "achieved": round(byts / t / 1e6, 1), "peak": hbm_peak, "unit": "GB/s",
                                 "frac": round(byts / t / 1e6 / hbm_peak, 4), "traffic": None,
                                 "kernel": "gbmm_bb (tensor-core tile kernel)", "also_fp64_tensor_frac": round(flops / t / 1e9 / tpeak, 4),
                                 "fp64_tensor_peak_TFLOPs": tpeak}})
    else:
        from bandedmatrices_b200.sharded import ShardedGbmm, shard_bounds

        j0, j1 = shard_bounds(n, rank, world)
        op = ShardedGbmm(n, (Al, Au), (Bl, Bu), j0, j1, A.data[max(0, j0 - Bu): min(n, j1 + Bl)])
        Cl = torch.empty((j1 - j0, Cm.data.shape[1]), dtype=torch.float64, device="cuda")
        Bloc = B.data[j0:j1]
        t, _ = _best(lambda: op(1.0, Bloc, 0.0, Cl), reps=5)
        bm.mul_(Cm, A, B)  # the unsharded product on this rank: the shard must reproduce its columns bit for bit
        same = bool(torch.equal(Cl.view(torch.int64), Cm.data[j0:j1].view(torch.int64)))
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ok = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        t = float(tt.item())
        out.update({"ms": round(t, 3), "GFLOPs": round(flops / t / 1e6, 1), "GBs": round(byts / t / 1e6, 1),
                    "sharding": f"columns of B and C over {world} ranks, static A halo of Bl+Bu columns, no exchange per product",
                    "sharded_bit_identical": bool(ok.item()),
                    "roofline": {"bound": "hbm", "achieved": round(byts / t / 1e6 / world, 1), "peak": hbm_peak, "unit": "GB/s per GPU",
                                 "frac": round(byts / t / 1e6 / world / hbm_peak, 4), "traffic": None}})
    if L is not None and rank == 0:
        # ---- parity at the FULL size: windows of columns replayed as principal sub-blocks (interior columns of a window see
        # exactly the FMAs of the full product: column j of C only touches A's columns [j-Bu, j+Bl] and B's column j) ----
        Wn, margin = 2048, 160
        starts = sorted({0, n // 7, n // 3, n // 2 + 333, (5 * n) // 6, n - Wn})
        checked, bad, maxerr = 0, 0, 0.0
        L.drv_set_threads(1)
        for j0 in starts:
            a_h, b_h = _host_band(A.data[j0:j0 + Wn]), _host_band(B.data[j0:j0 + Wn])
            c_h = np.zeros((Cm.data.shape[1], Wn), order="F")
            L.drv_gbmm(Wn, Wn, Wn, Al, Au, Bl, Bu, Al + Bl, Au + Bu, 1.0, a_h.ctypes.data, a_h.shape[0], b_h.ctypes.data, b_h.shape[0],
                       0.0, c_h.ctypes.data, c_h.shape[0])
            lo = 0 if j0 == 0 else margin
            hi = Wn if j0 + Wn == n else Wn - margin
            got = _host_band(Cm.data[j0 + lo:j0 + hi])
            ref = c_h[:, lo:hi]
            rr, jj = np.meshgrid(np.arange(ref.shape[0]), np.arange(j0 + lo, j0 + hi), indexing="ij")
            inm = (jj + rr - (Au + Bu) >= 0) & (jj + rr - (Au + Bu) < n)
            checked += int(inm.sum())
            bad += int((got[inm] != ref[inm]).sum())
            maxerr = max(maxerr, float(np.max(np.abs(got[inm] - ref[inm]) / np.maximum(1.0, np.abs(ref[inm])))))
        out["parity"] = {"entries_checked": checked, "columns": f"{len(starts)} windows of {Wn} columns incl. both matrix ends, n=2^{int(np.log2(n))}",
                         "bit_identical": bad == 0, "mismatching_entries": bad, "max_rel_err": maxerr, "tolerance": 1e-13,
                         "against": "per-column OpenBLAS dgbmv_ replay of _gbmm! (gbmm.jl:306-339) on principal sub-blocks"}
        # ---- CPU baseline: the same replay timed on a 2^18-column sample ----
        ns = min(n, 1 << 18)
        a_h, b_h = _host_band(A.data[:ns]), _host_band(B.data[:ns])
        c_h = np.zeros((Cm.data.shape[1], ns), order="F")
        tc = L.drv_gbmm(ns, ns, ns, Al, Au, Bl, Bu, Al + Bl, Au + Bu, 1.0, a_h.ctypes.data, a_h.shape[0], b_h.ctypes.data, b_h.shape[0], 0.0,
                        c_h.ctypes.data, c_h.shape[0])
        out["cpu_baseline"] = {"value": round(flops / (tc * n / ns) / 1e9, 2), "unit": "GFLOP/s", "ms_full_extrapolated": round(1e3 * tc * n / ns, 1),
                               "cores": 1, "kind": "reference",
                               "sample": f"first 2^{int(np.log2(ns))} of 2^{int(np.log2(n))} columns: _gbmm! = one dgbmv_64_ per column, replayed from C; linear in n"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def run_c4(bm, L, hbm_peak, rank=0, world=1, n=1 << 20, nrhs=256):
    import torch.distributed as dist

    l = u = 16
    ldab = 2 * l + u + 1
    A = bm.brand(n, n, l, u, seed=4)
    hd = bm.handle(torch.cuda.current_device())
    Wm = bm.BandedMatrix.undef((n, n), (l, 2 * l + u - l))
    keep = {}

    def widen():
        hd.check(hd.lib.bmb200_dband_widen(hd.h, n, l, u, vp(A.ptr), A.lda, vp(Wm.ptr), Wm.lda), "widen")

    def factor():
        keep["F"] = bm.lu_(Wm)

    out = {"workload": f"banded LU + solve Float64 n=2^{int(np.log2(n))}, (l,u)=(16,16), {nrhs} RHS (gbtrf! then gbtrs!)"}
    t_b = None
    if world == 1 or rank == 0:
        t_f, _ = _best(factor, reps=2, setup=widen)
        out.update({"gbtrf_ms": round(t_f, 2), "gbtrf_ns_per_column": round(1e6 * t_f / n, 1),
                    "nontrivial_pivots": int((keep["F"].ipiv != np.arange(1, n + 1)).sum())})
    q0, q1 = 0, nrhs
    if world > 1:
        from bandedmatrices_b200.sharded import ShardedSolve, rhs_bounds

        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        a.record()
        S = ShardedSolve(keep.get("F"), n, l, u, rank, world)
        b.record()
        b.synchronize()
        t_b = a.elapsed_time(b)
        keep["F"] = S.F
        q0, q1 = rhs_bounds(nrhs, rank, world)
    F = keep["F"]
    g = torch.Generator(device="cuda").manual_seed(77)
    Bfull = torch.rand((nrhs, n), dtype=torch.float64, device="cuda", generator=g)  # same on every rank
    Bm = Bfull[q0:q1].T            # this rank's column block, column-major n x (q1-q0)
    X = bm.colmajor(n, q1 - q0)
    t_s, _ = _best(lambda: bm.ldiv_(F, X), reps=3, setup=lambda: X.copy_(Bm))
    if world > 1:
        tt = torch.tensor([t_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_s = float(tt.item())
    fl_s = nrhs * n * (2.0 * l + 2.0 * (l + u) + 1)
    by_s = 2 * 8.0 * n * nrhs + world * (8.0 * ldab * n + 8.0 * n)  # every rank streams the factors once
    out.update({"gbtrs_ms": round(t_s, 2), "gbtrs_GFLOPs": round(fl_s / t_s / 1e6, 1), "gbtrs_GBs": round(by_s / t_s / 1e6, 1),
                "gbtrs_algorithmic_bytes": by_s,
                "roofline": {"bound": "hbm", "achieved": round(by_s / t_s / 1e6 / world, 1), "peak": hbm_peak, "unit": "GB/s" + (" per GPU" if world > 1 else ""),
                             "frac": round(by_s / t_s / 1e6 / world / hbm_peak, 4), "traffic": None, "kernel": "gbtrs_slot (forward + backward sweep)",
                             "note": "the sweeps are chains of n dependent steps per right-hand side: latency-bound far below the byte roofline (DESIGN.md)"}})
    if world > 1:
        out["sharding"] = f"right-hand sides over {world} ranks ({q1 - q0} columns each), factors broadcast once ({round(t_b, 1)} ms incl. allocation), no exchange during the solve"
    if L is not None:
        # ---- parity at the FULL size against OpenBLAS dgbtrf_64_ / dgbtrs_64_ on the same inputs ----
        a_h = _host_band(A.data)
        ab = np.zeros((ldab, n), order="F")
        ab[l:, :] = a_h
        ipiv = np.zeros(n, dtype=np.int64)
        info = np.zeros(1, dtype=np.int64)
        L.drv_set_threads(os.cpu_count() or 1)
        tcf = L.drv_gbtrf(n, n, l, u, ab.ctypes.data, ldab, ipiv.ctypes.data, info.ctypes.data)
        nchk = min(16 if world == 1 else 2, q1 - q0)
        bh = np.asfortranarray(Bm[:, :nchk].cpu().numpy())
        tcs = L.drv_gbtrs(n, l, u, nchk, ab.ctypes.data, ldab, ipiv.ctypes.data, bh.ctypes.data, n, info.ctypes.data)
        par = {"ipiv_bit_identical": _bits_equal(np.asarray(F.ipiv), ipiv),
               "factors_bit_identical": _bits_equal(_host_band(F.factors.data), ab),
               f"solution_bit_identical_{nchk}rhs": _bits_equal(X[:, :nchk].cpu().numpy(), bh),
               "against": "OpenBLAS 0.3.30 dgbtrf_64_ / dgbtrs_64_ at the full n"}
        if world > 1:
            ok = torch.tensor([1 if all(v for k, v in par.items() if k != "against") else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            par["all_ranks"] = bool(ok.item())
        # residual of the first column (north_star: relative residual <= 1e-12 * cond)
        r = Bm[:, 0].clone()
        bm.mul_(r, A, X[:, 0].contiguous(), -1.0, 1.0)
        par["rel_residual_col0"] = float(r.abs().max() / (X[:, 0].abs().max() * (l + u + 1)))
        out["parity"] = par
        if rank == 0:
            out["cpu_baseline"] = {"gbtrf_ms": round(1e3 * tcf, 1), "gbtrf_ns_per_column": round(1e9 * tcf / n, 1),
                                   "gbtrs_ms_extrapolated": round(1e3 * tcs * nrhs / nchk, 1),
                                   "gbtrs_GFLOPs": round(fl_s / (tcs * nrhs / nchk) / 1e9, 2), "cores": os.cpu_count(), "kind": "reference",
                                   "sample": f"dgbtrf_64_ on the full matrix; dgbtrs_64_ on {nchk} of {nrhs} right-hand sides, linear in nrhs"}
    return out


def run_c4_cholesky(bm, L, n=1 << 20, kd=16, nrhs=256):
    """The C4 shape through the banded Cholesky: pbtrf!('U', n=2^20, kd=16) + pbtrs! with 256 right-hand sides, on an SPD band
    (random band + dominant diagonal).  kd <= 64 is the DPBTF2 regime: the factor is compared BIT FOR BIT with OpenBLAS
    dpbtrf_64_ at the full size."""
    out = {"workload": f"banded Cholesky + solve Float64 n=2^{int(np.log2(n))}, kd={kd}, {nrhs} RHS (pbtrf! then pbtrs!, uplo 'U')"}
    g = torch.Generator(device="cuda").manual_seed(99)
    U0 = torch.rand((n, kd + 1), dtype=torch.float64, device="cuda", generator=g) - 0.5
    U0[:, kd] = 2.0 * (kd + 1) + torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    U = U0.clone()
    keep = {}

    def factor():
        keep["info"] = bm.pbtrf_("U", n, kd, U)[1]

    t_f, _ = _best(factor, reps=2, setup=lambda: U.copy_(U0))
    Bfull = torch.rand((nrhs, n), dtype=torch.float64, device="cuda", generator=g)
    X = bm.colmajor(n, nrhs)
    t_s, _ = _best(lambda: bm.pbtrs_("U", n, kd, U, X), reps=3, setup=lambda: X.copy_(Bfull.T))
    out.update({"info": int(keep["info"]), "pbtrf_ms": round(t_f, 2), "pbtrf_ns_per_column": round(1e6 * t_f / n, 1), "pbtrs_ms": round(t_s, 2),
                "note": "both are chains of n dependent steps (sqrt + divide per column; a sweep per right-hand side): latency-bound, no roofline claim"})
    if L is not None:
        import oracle

        ob = oracle.backend("OB")
        ab = np.asfortranarray(_host_band(U0))
        ob.set_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        iref = ob.pbtrf("U", n, kd, ab, kd + 1)
        tcf = time.perf_counter() - t0
        nchk = 16
        bh = np.asfortranarray(Bfull[:nchk].T.cpu().numpy())
        t0 = time.perf_counter()
        ob.pbtrs("U", n, kd, nchk, ab, kd + 1, bh, n)
        tcs = time.perf_counter() - t0
        ob.set_threads(1)
        got = X[:, :nchk].cpu().numpy()
        out["parity"] = {"info_equal": bool(iref == keep["info"] == 0), "factor_bit_identical": _bits_equal(_host_band(U), ab),
                         f"solution_max_rel_diff_{nchk}rhs": float(np.max(np.abs(got - bh)) / np.max(np.abs(bh))), "solution_tolerance": 1e-13,
                         "against": "OpenBLAS 0.3.30 dpbtrf_64_ (= DPBTF2 for kd <= 64) / dpbtrs_64_ at the full n"}
        out["cpu_baseline"] = {"pbtrf_ms": round(1e3 * tcf, 1), "pbtrf_ns_per_column": round(1e9 * tcf / n, 1),
                               "pbtrs_ms_extrapolated": round(1e3 * tcs * nrhs / nchk, 1), "cores": os.cpu_count(), "kind": "reference",
                               "sample": f"dpbtrf_64_ on the full matrix; dpbtrs_64_ on {nchk} of {nrhs} right-hand sides, linear in nrhs"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def run_wide_gbmm(bm, L, n=1 << 16, l=1024):
    """Banded x banded with C5-sized bands: (l,l) x (l,l) -> (2l,2l), the K-blocked tensor-core kernel (gbmm_wide.cu).  Timed at
    n = 2^16; verified bit for bit against the per-column dgbmv_ replay on a smaller principal block (the replay costs
    4 l^2 flops per column on one CPU core)."""
    tpeak, tsrc = fp64_peak()
    out = {"workload": f"BandedMatrix*BandedMatrix Float64 n=2^{int(np.log2(n))}, ({l},{l})x({l},{l})->({2*l},{2*l})"}
    A, B = bm.brand(n, n, l, l, seed=12), bm.brand(n, n, l, l, seed=13)
    Cm = bm.BandedMatrix.undef((n, n), (2 * l, 2 * l))
    flops = 2.0 * (2 * l + 1) ** 2 * n
    t, tm = _best(lambda: bm.mul_(Cm, A, B), reps=3)
    out.update({"ms": round(t, 2), "TFLOPs": round(flops / t / 1e9, 2), "kernel_path": int(bm.handle(0).last_gbmm_path()),
                "roofline": {"bound": "tensor", "achieved": round(flops / t / 1e9, 2), "peak": tpeak, "unit": "TFLOP/s (FP64)",
                             "frac": round(flops / t / 1e9 / tpeak, 4), "traffic": None, "peak_source": tsrc,
                             "kernel": "gbmm_bb_kblock (64x64 C tiles, 32-deep K-blocks, DMMA.8x8x4)"}})
    if L is not None:
        ns = 4 * l + 512
        a_h, b_h = _host_band(A.data[:ns]), _host_band(B.data[:ns])
        c_h = np.zeros((4 * l + 1, ns), order="F")
        L.drv_set_threads(1)
        tc = L.drv_gbmm(ns, ns, ns, l, l, l, l, 2 * l, 2 * l, 1.0, a_h.ctypes.data, a_h.shape[0], b_h.ctypes.data, b_h.shape[0], 0.0,
                        c_h.ctypes.data, c_h.shape[0])
        As, Bs = bm.BandedMatrix(A.data[:ns].clone(), ns, l, l), bm.BandedMatrix(B.data[:ns].clone(), ns, l, l)
        Cs = bm.BandedMatrix.undef((ns, ns), (2 * l, 2 * l))
        bm.mul_(Cs, As, Bs)
        got = _host_band(Cs.data)
        rr, jj = np.meshgrid(np.arange(got.shape[0]), np.arange(ns), indexing="ij")
        inm = (jj + rr - 2 * l >= 0) & (jj + rr - 2 * l < ns)
        out["parity"] = {"principal_block": ns, "entries_checked": int(inm.sum()), "bit_identical": bool(np.array_equal(got[inm], c_h[inm])),
                         "against": "per-column OpenBLAS dgbmv_ replay of _gbmm! (gbmm.jl:306-339)"}
        out["cpu_baseline"] = {"ms_full_extrapolated": round(1e3 * tc * n / ns, 0), "cores": 1, "kind": "reference",
                               "sample": f"leading {ns} columns, one dgbmv_64_ per column; linear in n away from the ends"}
    return out


def run_c5(bm, L, hbm_peak, N=1024):
    n = N * N
    tpeak, tsrc = fp64_peak()
    out = {"workload": f"2-D finite-difference Laplacian N={N}: n=2^{int(np.log2(n))}, l=u={N}, lu + ldiv! (1 RHS)"}
    # ---- small leading block first: warm-up of every kernel involved AND the comparison with OpenBLAS' blocked DGBTRF ----
    ns = 16 * N
    As = laplacian(bm, N, ns)
    Fs = bm.lu(As)
    bs = torch.ones(ns, dtype=torch.float64, device="cuda")
    xs = bs.clone()
    bm.ldiv_(Fs, xs)
    par = {}
    if L is not None:
        ldab = 3 * N + 1
        ab = np.zeros((ldab, ns), order="F")
        ab[N:, :] = _host_band(As.data)
        ipiv = np.zeros(ns, dtype=np.int64)
        info = np.zeros(1, dtype=np.int64)
        L.drv_set_threads(os.cpu_count() or 1)
        tcf = L.drv_gbtrf(ns, ns, N, N, ab.ctypes.data, ldab, ipiv.ctypes.data, info.ctypes.data)
        xh = np.ones(ns)
        tcs = L.drv_gbtrs(ns, N, N, 1, ab.ctypes.data, ldab, ipiv.ctypes.data, xh.ctypes.data, ns, info.ctypes.data)
        got = _host_band(Fs.factors.data)
        par["sample_n"] = ns
        par["sample_ipiv_bit_identical"] = _bits_equal(np.asarray(Fs.ipiv), ipiv)
        par["sample_factors_max_abs_diff"] = float(np.max(np.abs(got - ab)))   # OpenBLAS' blocked DGBTRF rounds its DGEMMs differently
        par["sample_solution_max_rel_diff"] = float(np.max(np.abs(xs.cpu().numpy() - xh)) / np.max(np.abs(xh)))
        par["against"] = "OpenBLAS 0.3.30 dgbtrf_64_ (blocked) / dgbtrs_64_ on the leading 2^14 columns"
        flops_s = 2.0 * ns * N * N + ns * N
        out["cpu_baseline"] = {"lu_ms_extrapolated": round(1e3 * tcf * n / ns, 0), "lu_TFLOPs": round(flops_s / tcf / 1e12, 3),
                               "solve_ms_extrapolated": round(1e3 * tcs * n / ns, 0), "cores": os.cpu_count(), "kind": "reference",
                               "sample": f"leading 2^{int(np.log2(ns))} of 2^{int(np.log2(n))} columns (26 GB of host memory for the full factor storage); work per column is constant"}
    del As, Fs, xs, bs
    torch.cuda.empty_cache()
    # ---- the full problem ----
    A = laplacian(bm, N)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    F = bm.lu(A)
    b.record()
    b.synchronize()
    t_f = a.elapsed_time(b)
    rhs = torch.ones(n, dtype=torch.float64, device="cuda")
    x = rhs.clone()
    a.record()
    bm.ldiv_(F, x)
    b.record()
    b.synchronize()
    t_s = a.elapsed_time(b)
    r = rhs.clone()
    bm.mul_(r, A, x, -1.0, 1.0)
    flops = 2.0 * n * N * N + n * N
    par["pivots_identity_full"] = bool((torch.as_tensor(F.ipiv) == torch.arange(1, n + 1)).all())
    par["max_residual_over_max_x_full"] = float(r.abs().max() / x.abs().max())
    out.update({"lu_ms_incl_widen": round(t_f, 1), "lu_TFLOPs": round(flops / t_f / 1e9, 3), "solve_ms": round(t_s, 1),
                "flops": flops,
                "roofline": {"bound": "tensor", "achieved": round(flops / t_f / 1e9, 3), "peak": tpeak, "unit": "TFLOP/s (FP64)",
                             "frac": round(flops / t_f / 1e9 / tpeak, 4), "traffic": None, "peak_source": tsrc,
                             "kernel": "gbtrf_strip_kernel (strip-resident LU: diagonal-block chain + DMMA strip updates, verified diagonal pivots)"},
                "parity": par})
    del F, x, r
    torch.cuda.empty_cache()
    # ---- the same system through the banded Cholesky: cholesky(Symmetric(A)) \\ b (pbtrf! / pbtrs!, BandedCholesky.jl) ----
    try:
        out["cholesky"] = _c5_cholesky(bm, L, A, rhs, N, tpeak, tsrc)
    except Exception as e:  # noqa: BLE001
        out["cholesky"] = {"error": repr(e)}
    del A, rhs
    torch.cuda.empty_cache()
    return out


def _c5_cholesky(bm, L, A, rhs, N, tpeak, tsrc):
    n = A.n
    res = {"workload": f"cholesky(Symmetric(A)) and its ldiv! on the same Laplacian: pbtrf!('U', n=2^{int(np.log2(n))}, kd={N}) + pbtrs! (1 RHS)"}
    U = A.data[:, : N + 1].clone()  # the stored 'U' triangle (band rows 0..N), (n, N+1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _, info = bm.pbtrf_("U", n, N, U)
    b.record()
    b.synchronize()
    t_f = a.elapsed_time(b)
    x = rhs.clone()
    bm.pbtrs_("U", n, N, U, x)  # warm-up (workspace for the transposed factor)
    x = rhs.clone()
    a.record()
    bm.pbtrs_("U", n, N, U, x)
    b.record()
    b.synchronize()
    t_s = a.elapsed_time(b)
    r = rhs.clone()
    bm.mul_(r, A, x, -1.0, 1.0)
    flops = float(n) * (N + 1.0) ** 2
    res.update({"info": int(info), "pbtrf_ms": round(t_f, 1), "pbtrf_TFLOPs": round(flops / t_f / 1e9, 3), "pbtrs_ms": round(t_s, 1), "flops": flops,
                "roofline": {"bound": "tensor", "achieved": round(flops / t_f / 1e9, 3), "peak": tpeak, "unit": "TFLOP/s (FP64)",
                             "frac": round(flops / t_f / 1e9 / tpeak, 4), "traffic": None, "peak_source": tsrc,
                             "kernel": "pb_potf2_reg (factorisation + substitution) + pb_syrk (DMMA) per 64-column panel, CUDA graph"},
                "parity": {"max_residual_over_max_x_full": float(r.abs().max() / x.abs().max())}})
    if L is not None:
        import oracle

        ob = oracle.backend("OB")
        ns = 16 * N
        ab = np.asfortranarray(_host_band(A.data[:ns, : N + 1]))
        ob.set_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        iref = ob.pbtrf("U", ns, N, ab, N + 1)
        tcf = time.perf_counter() - t0
        xh = np.ones((ns, 1), order="F")
        t0 = time.perf_counter()
        ob.pbtrs("U", ns, N, 1, ab, N + 1, xh, ns)
        tcs = time.perf_counter() - t0
        ob.set_threads(1)
        Us = A.data[:ns, : N + 1].clone()
        _, i2 = bm.pbtrf_("U", ns, N, Us)
        got = _host_band(Us)
        rr, jj = np.meshgrid(np.arange(N + 1), np.arange(ns), indexing="ij")
        inm = jj + rr - N >= 0
        res["parity"].update({"sample_n": ns, "sample_info_equal": bool(iref == i2 == 0),
                              "sample_factor_max_rel_diff": float(np.max(np.abs(got[inm] - ab[inm])) / np.max(np.abs(ab[inm]))),
                              "tolerance": 1e-12, "against": "OpenBLAS 0.3.30 dpbtrf_64_ (blocked DPBTRF) on the leading 2^14 columns"})
        res["cpu_baseline"] = {"pbtrf_ms_extrapolated": round(1e3 * tcf * n / ns, 0), "pbtrs_ms_extrapolated": round(1e3 * tcs * n / ns, 0),
                               "cores": os.cpu_count(), "kind": "reference", "sample": f"leading 2^{int(np.log2(ns))} of 2^{int(np.log2(n))} columns; work per column is constant"}
    return res


def run_configs(bm, L, hbm_peak, rank=0, world=1, skip=()):
    """All of the above; failures are reported per entry, never raised (the headline line must survive)."""
    out = {}
    plan = [("C3", lambda: run_c3(bm, L, hbm_peak, rank, world)), ("C4", lambda: run_c4(bm, L, hbm_peak, rank, world))]
    if world == 1:
        plan = [("C1", lambda: run_c1(bm, L, hbm_peak))] + plan + [("C4_chol", lambda: run_c4_cholesky(bm, L)), ("C3_wide", lambda: run_wide_gbmm(bm, L)),
                                                                   ("C5", lambda: run_c5(bm, L, hbm_peak))]
    for name, fn in plan:
        if name in skip:
            continue
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)}
        out[name]["wall_s"] = round(time.perf_counter() - t0, 1)
        torch.cuda.empty_cache()
    return out
