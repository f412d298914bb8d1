"""Secondary kernels timed on one GPU (reported under "extras" by bench.py --extras: triangular band tbsv/tbmv, symmetric band
sbmv, band-aligned axpy), and `cpu_extras`: the reference's CPU timings of the BASELINE configs on bounded samples, which the
reference arm (bench.py --impl reference) reports under "configs".  The GPU side of C1/C3/C4/C5 lives in bench_configs.py.
CUDA events on the launching stream, best of `reps` after one warm-up."""
import numpy as np
import torch


def _time(fn, reps=5, setup=None):
    best = 1e30
    for i in range(reps + 1):
        if setup:
            setup()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        b.synchronize()
        if i:
            best = min(best, a.elapsed_time(b))
    return best


def cpu_extras(L, cores):
    """The reference's CPU path for C1/C3/C4/C5 beside the GPU numbers: OpenBLAS 0.3.30 entered through the
    Fortran symbols the reference ccalls, replaying its call sequences from C (oracle/blasdriver.c) on BOUNDED samples of
    each workload; every entry states its sample and the linear extrapolation to the full size."""
    import ctypes as C

    i64, dbl, vp = C.c_int64, C.c_double, C.c_void_p
    L.drv_gbmv.restype = dbl
    L.drv_gbmv.argtypes = [i64] * 4 + [dbl, vp, i64, vp, dbl, vp]
    L.drv_gbmm.restype = dbl
    L.drv_gbmm.argtypes = [i64] * 9 + [dbl, vp, i64, vp, i64, dbl, vp, i64]
    L.drv_gbtrf.restype = dbl
    L.drv_gbtrf.argtypes = [i64] * 4 + [vp, i64, vp, vp]
    L.drv_gbtrs.restype = dbl
    L.drv_gbtrs.argtypes = [i64] * 4 + [vp, i64, vp, vp, i64, vp]
    rng = np.random.default_rng(7)
    out = {"cores": cores, "library": "OpenBLAS 0.3.30 ILP64 (numpy's), dgbmv pinned to 1 thread, LAPACK calls on all cores"}

    def lu_solve(n, l, u, nrhs, dominant=False):
        ldab = 2 * l + u + 1
        ab = np.zeros((ldab, n), order="F")
        ab[l:, :] = rng.random((l + u + 1, n))
        if dominant:
            ab[l + u, :] += 2.0 * (l + u + 1)
        ipiv = np.zeros(n, dtype=np.int64)
        info = np.zeros(1, dtype=np.int64)
        L.drv_set_threads(cores)
        tf = L.drv_gbtrf(n, n, l, u, ab.ctypes.data, ldab, ipiv.ctypes.data, info.ctypes.data)
        b = np.asfortranarray(rng.random((n, nrhs)))
        ts = L.drv_gbtrs(n, l, u, nrhs, ab.ctypes.data, ldab, ipiv.ctypes.data, b.ctypes.data, n, info.ctypes.data)
        return 1e3 * tf, 1e3 * ts

    # C1: the README workload in full
    n = 10000
    a = np.asfortranarray(rng.random((8, n)))
    x, y = rng.standard_normal(n), np.zeros(n)
    L.drv_set_threads(1)
    t_ab = min(L.drv_gbmv(n, n, 4, 3, 1.0, a.ctypes.data, 8, x.ctypes.data, 0.0, y.ctypes.data) for _ in range(5))
    c = np.zeros((15, n), order="F")
    t_aa = min(L.drv_gbmm(n, n, n, 4, 3, 4, 3, 8, 6, 1.0, a.ctypes.data, 8, a.ctypes.data, 8, 0.0, c.ctypes.data, 15) for _ in range(3))
    tf, ts = lu_solve(n, 4, 3, 1)
    out["C1"] = {"Ab_us": round(1e6 * t_ab, 1), "AA_us": round(1e6 * t_aa, 1), "solve_us": round(1e3 * (tf + ts), 1), "sample": "full size"}
    # C3: _gbmm! = one dgbmv_ per column of C (gbmm.jl:306-339); sample of 2^18 columns, linear in n
    ns, full = 1 << 18, 1 << 22
    a = np.asfortranarray(rng.random((65, ns)))
    b = np.asfortranarray(rng.random((65, ns)))
    c = np.zeros((129, ns), order="F")
    L.drv_set_threads(1)
    t = L.drv_gbmm(ns, ns, ns, 32, 32, 32, 32, 64, 64, 1.0, a.ctypes.data, 65, b.ctypes.data, 65, 0.0, c.ctypes.data, 129)
    out["C3"] = {"sample": "n=2^18 of 2^22 columns", "sample_ms": round(1e3 * t, 1), "full_ms_extrapolated": round(1e3 * t * full / ns, 1)}
    del a, b, c
    # C4: dgbtrf on the full matrix; dgbtrs on 16 of the 256 right-hand sides, linear in nrhs
    tf, ts = lu_solve(1 << 20, 16, 16, 16)
    out["C4"] = {"sample": "n=2^20 in full; 16 of 256 RHS for dgbtrs", "gbtrf_ms": round(tf, 1), "gbtrs_sample_ms": round(ts, 1),
                 "gbtrs_full_ms_extrapolated": round(ts * 256 / 16, 1)}
    # C5: n=2^20, l=u=1024 needs 26 GB of host memory and ~30 s; sample n=2^14, linear in n (work per column is constant)
    ns, full = 1 << 14, 1 << 20
    tf, ts = lu_solve(ns, 1024, 1024, 1, dominant=True)
    out["C5"] = {"sample": "n=2^14 of 2^20 columns, l=u=1024, diagonally dominant, 1 RHS", "lu_sample_ms": round(tf, 1), "solve_sample_ms": round(ts, 1),
                 "lu_full_ms_extrapolated": round(tf * full / ns, 1), "solve_full_ms_extrapolated": round(ts * full / ns, 1)}
    # symmetric band: dsbmv_ 'U', k=3 on n=2^24 of 2^27 rows (1 thread), linear in n
    ns, full = 1 << 24, 1 << 27
    a = np.asfortranarray(rng.random((4, ns)))
    xv, yv = rng.random(ns), np.zeros(ns)
    L.drv_sbmv.restype = dbl
    L.drv_sbmv.argtypes = [C.c_char, i64, i64, dbl, vp, i64, vp, dbl, vp]
    L.drv_set_threads(1)
    t_sb = min(L.drv_sbmv(b"U", ns, 3, 1.0, a.ctypes.data, 4, xv.ctypes.data, 0.0, yv.ctypes.data) for _ in range(3))
    out["SB"] = {"sample": "n=2^24 of 2^27 rows, k=3, 1 thread", "sbmv_sample_ms": round(1e3 * t_sb, 1), "sbmv_full_ms_extrapolated": round(1e3 * t_sb * full / ns, 1),
                 "GBs": round(8.0 * ns * 6 / t_sb / 1e9, 2)}
    del a, xv, yv
    # triangular band: dtbsv_/dtbmv_ 'U','N','N', k=1024 on n=2^16 of 2^20 columns (1 thread: OpenBLAS' level-2 band
    # kernels are not threaded at these sizes), linear in n
    ns, full, k = 1 << 16, 1 << 20, 1024
    a = np.asfortranarray(rng.random((k + 1, ns))) / (2 * k)
    a[k, :] = 2.0
    xv = np.ones(ns)
    L.drv_tb.restype = dbl
    L.drv_tb.argtypes = [C.c_int, C.c_char, C.c_char, i64, i64, vp, i64, vp]
    L.drv_set_threads(1)
    t_sv = L.drv_tb(0, b"U", b"N", ns, k, a.ctypes.data, k + 1, xv.ctypes.data)
    t_mv = L.drv_tb(1, b"U", b"N", ns, k, a.ctypes.data, k + 1, xv.ctypes.data)
    out["TB"] = {"sample": "n=2^16 of 2^20 columns, k=1024, 1 thread", "tbsv_sample_ms": round(1e3 * t_sv, 1), "tbmv_sample_ms": round(1e3 * t_mv, 1),
                 "tbsv_full_ms_extrapolated": round(1e3 * t_sv * full / ns, 1), "tbmv_full_ms_extrapolated": round(1e3 * t_mv * full / ns, 1)}
    return out


def run_secondary(bm):
    """Secondary kernels: triangular band solve / multiply, symmetric band matvec, band-aligned axpy."""
    from bench_configs import laplacian

    out = {}
    # ---- triangular band: the Gauss-Seidel half of examples/finitedifference_2d.jl:18-26 at C5 size ----
    N = 1024
    A = laplacian(bm, N)
    n = N * N
    b = torch.ones(n, dtype=torch.float64, device="cuda")
    x = b.clone()
    t_sv = _time(lambda: bm.ldiv_tri_("U", False, A, x), reps=2, setup=lambda: x.copy_(b))
    t_mv = _time(lambda: bm.lmul_tri_("U", False, A, x), reps=2, setup=lambda: x.copy_(b))
    by = 8.0 * n * (N + 1) + 16.0 * n
    # ---- symmetric band matvec: the C2 shape with only the upper triangle stored ----
    ns, ks = 1 << 27, 3
    Sd = torch.rand((ns, ks + 1), dtype=torch.float64, device="cuda")
    xs, ys = torch.rand(ns, dtype=torch.float64, device="cuda"), torch.empty(ns, dtype=torch.float64, device="cuda")
    t_sb = _time(lambda: bm.sbmv_("U", ks, 1.0, Sd, xs, 0.0, ys), reps=5)
    by_sb = 8.0 * ns * (ks + 1 + 2)
    out["SB"] = {"n": ns, "k": ks, "sbmv_U_ms": round(t_sb, 3), "GBs": round(by_sb / t_sb / 1e6, 1), "algorithmic_bytes": by_sb}
    del Sd, xs, ys
    # ---- band-aligned elementwise ops: Y += a*X on the C2 shape, equal and unequal bandwidths ----
    ne = 1 << 26
    Xe, Ye = bm.brand(ne, ne, 4, 3, seed=11), bm.brand(ne, ne, 4, 3, seed=12)
    t_eq = _time(lambda: bm.axpy_(0.5, Xe, Ye), reps=5)
    Yw = bm.brand(ne, ne, 5, 4, seed=13)
    t_ne = _time(lambda: bm.axpy_(0.5, Xe, Yw), reps=3)
    out["EW"] = {"n": ne, "axpy_equal_bands_ms": round(t_eq, 3), "axpy_equal_GBs": round(3 * 8.0 * 8 * ne / t_eq / 1e6, 1),
                 "axpy_(4,3)_into_(5,4)_ms": round(t_ne, 3), "note": "unequal bandwidths also run the BandError counting pass over X and synchronise"}
    del Xe, Ye, Yw
    # ---- gbmv on the wide band (the C5 residual b - A x: every gbmv with l+u+1 > 16 takes the sweep kernel) ----
    xw, yw = torch.rand(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
    t_gw = _time(lambda: bm.mul_(yw, A, xw, 1.0, 0.0), reps=3)
    by_gw = 8.0 * n * (2 * N + 1) + 16.0 * n
    out["GBMV_wide"] = {"n": n, "l": N, "u": N, "ms": round(t_gw, 3), "GBs": round(by_gw / t_gw / 1e6, 1), "algorithmic_bytes": by_gw}
    del xw, yw
    # ---- the other element types: gbmv 'N' and 'C' on the C2 band shape (4,3), n = 2^26 ----
    nt = 1 << 26
    ty = {}
    for name, dt, eb in (("Float32", torch.float32, 4), ("ComplexF32", torch.complex64, 8), ("ComplexF64", torch.complex128, 16)):
        Ad = torch.randn((nt, 8), dtype=dt, device="cuda")
        xt, yt = torch.randn(nt, dtype=dt, device="cuda"), torch.empty(nt, dtype=dt, device="cuda")
        tn = _time(lambda: bm.gbmv_("N", nt, 4, 3, 1.0, Ad, xt, 0.0, yt), reps=3)
        tc = _time(lambda: bm.gbmv_("C", nt, 4, 3, 1.0, Ad, xt, 0.0, yt), reps=3)
        byt = float(eb) * nt * 10
        ty[name] = {"gbmv_N_ms": round(tn, 3), "gbmv_N_GBs": round(byt / tn / 1e6, 1), "gbmv_C_ms": round(tc, 3), "gbmv_C_GBs": round(byt / tc / 1e6, 1)}
        del Ad, xt, yt
    out["TYPED"] = {"n": nt, "bands": [4, 3], "algorithmic_bytes_per_element_size": 10 * nt, **ty}
    out["TB"] = {"n": n, "k": N, "tbsv_U_ms": round(t_sv, 2), "tbmv_U_ms": round(t_mv, 3), "tbmv_GBs": round(by / t_mv / 1e6, 1),
                 "algorithmic_bytes": by}
    return out
