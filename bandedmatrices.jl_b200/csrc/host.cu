// host.cu -- host-buffer entry points (what a Fortran-ABI caller holding HOST arrays gets).
// Inputs are streamed host->device in chunks on a copy stream while the previous chunk is being
// reduced on the compute stream; results are copied back before the call returns.  If the caller's
// arrays are page-locked the copies are true DMA at PCIe rate, otherwise CUDA stages them.
#include "common.cuh"

int bmb_gbmv_device(bmb200_ctx *h, bool tr, i64 m, i64 n, i64 kl, i64 ku, double alpha, const double *dA, i64 lda,
                    const double *dx, i64 incx, double beta, double *dy, i64 incy);

// Device arena carved from the grow-only scratch allocation.
struct Arena {
    char *base;
    size_t off = 0;
    explicit Arena(void *p) : base((char *)p) {}
    template <typename T>
    T *take(size_t count)
    {
        T *p = (T *)(base + off);
        off += ((count * sizeof(T) + 511) / 512) * 512;
        return p;
    }
};

extern "C" int bmb200_dgbmv_host(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
                                 double alpha, const double *hA, int64_t lda, const double *hx, int64_t incx,
                                 double beta, double *hy, int64_t incy)
{
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (kl < 0) return -5;
    if (ku < 0) return -6;
    if (lda < kl + ku + 1) return -9;
    if (incx == 0) return -11;
    if (incy == 0) return -14;
    if (m == 0 || n == 0) return 0;
    DeviceGuard g(h->device);
    const i64 lenx = tr ? m : n, leny = tr ? n : m;
    const i64 ax = incx < 0 ? -incx : incx, ay = incy < 0 ? -incy : incy;
    const i64 spanx = (lenx - 1) * ax + 1, spany = (leny - 1) * ay + 1;
    // BLAS convention for negative increments: element 0 sits at the END of the span
    const double *hxlow = hx;   // lowest address of the x span (== the pointer BLAS receives)
    double *hylow = hy;

    // chunk over the OUTPUT index (rows for 'N', columns for 'T'); both walk A's columns forward
    i64 chunk = (i64)(256ull << 20) / (lda * 8);   // ~256 MiB of A per chunk
    if (chunk < 1024) chunk = 1024;
    if (chunk > leny) chunk = leny;
    const size_t abuf = (size_t)(chunk + kl + ku + 2) * lda;
    size_t need = (2 * abuf + spanx + spany) * sizeof(double) + 4 * 512;
    int rc = bmb_ensure_scratch(h, need);
    if (rc) return rc;
    Arena ar(h->scratch);
    double *dAbuf[2] = {ar.take<double>(abuf), ar.take<double>(abuf)};
    double *dx = ar.take<double>(spanx);
    double *dy = ar.take<double>(spany);

    cudaStream_t cs = h->copy_stream, ks = h->stream;
    BMB_CUDA(h, cudaEventRecord(h->ev[3], ks));  // the copy stream must not overtake earlier users of the scratch
    BMB_CUDA(h, cudaStreamWaitEvent(cs, h->ev[3], 0));
    BMB_CUDA(h, cudaMemcpyAsync(dx, hxlow, spanx * sizeof(double), cudaMemcpyHostToDevice, cs));
    // y goes up when it is read (beta != 0) or when it is strided (the gaps must survive the round trip)
    if (beta != 0.0 || ay != 1)
        BMB_CUDA(h, cudaMemcpyAsync(dy, hylow, spany * sizeof(double), cudaMemcpyHostToDevice, cs));
    const double *x0 = incx > 0 ? dx : dx + (lenx - 1) * ax;  // element 0 on the device
    double *y0 = incy > 0 ? dy : dy + (leny - 1) * ay;

    // ev[0..1]: "A buffer b filled"; ev[2..3]: "A buffer b consumed"
    int b = 0;
    i64 ncopies = 0;
    bool xy_waited = false;
    for (i64 o0 = 0; o0 < leny; o0 += chunk) {
        const i64 o1 = imin64(o0 + chunk, leny);
        i64 c0, c1, kls, kus, ms, ns, xoff;
        if (!tr) {  // rows [o0,o1) need columns [o0-kl, o1+ku): S = A[o0:o1, c0:c1]
            c0 = imax64(0, o0 - kl);
            c1 = imin64(n, o1 + ku);
            kus = ku + (o0 - c0);
            kls = kl - (o0 - c0);
            ms = o1 - o0;
            ns = c1 - c0;
            xoff = c0;
        } else {  // columns [o0,o1) need rows [o0-ku, o1+kl) of x: S = A[r0:r1, o0:o1]
            c0 = o0;
            c1 = o1;
            const i64 r0 = imax64(0, o0 - ku), r1 = imin64(m, o1 + kl);
            kus = ku - (o0 - r0);
            kls = kl + (o0 - r0);
            ms = r1 - r0;
            ns = c1 - c0;
            xoff = r0;
        }
        double *ys = y0 + o0 * incy;
        if (ms <= 0 || ns <= 0) {  // nothing of A touches this output range: y <- beta*y
            if (beta != 1.0) {
                if (!xy_waited) {
                    BMB_CUDA(h, cudaEventRecord(h->ev[0], cs));
                    BMB_CUDA(h, cudaStreamWaitEvent(ks, h->ev[0], 0));
                    xy_waited = true;
                }
                double *ylow = incy > 0 ? ys : y0 + (o1 - 1) * incy;
                rc = bmb200_dfill_lmul(h, beta, ylow, o1 - o0, 1, 0, ay);
                if (rc) return rc;
            }
            continue;
        }
        if (ncopies >= 2) BMB_CUDA(h, cudaStreamWaitEvent(cs, h->ev[2 + b], 0));
        BMB_CUDA(h, cudaMemcpyAsync(dAbuf[b], hA + c0 * lda, (size_t)(c1 - c0) * lda * sizeof(double),
                                    cudaMemcpyHostToDevice, cs));
        BMB_CUDA(h, cudaEventRecord(h->ev[b], cs));
        BMB_CUDA(h, cudaStreamWaitEvent(ks, h->ev[b], 0));  // also orders the x / y uploads before the kernel
        xy_waited = true;
        rc = bmb_gbmv_device(h, tr, ms, ns, kls, kus, alpha, dAbuf[b], lda, x0 + xoff * incx, incx, beta, ys, incy);
        if (rc) return rc;
        BMB_CUDA(h, cudaEventRecord(h->ev[2 + b], ks));
        b ^= 1;
        ++ncopies;
    }
    BMB_CUDA(h, cudaMemcpyAsync(hylow, dy, spany * sizeof(double), cudaMemcpyDeviceToHost, ks));
    BMB_CUDA(h, cudaStreamSynchronize(ks));
    BMB_CUDA(h, cudaStreamSynchronize(cs));
    return 0;
}

// dgbtrf_ + dgbtrs_ on host arrays (LAPACK dgbsv semantics): hAB (ldab x n, LU storage) is overwritten by the
// factors, h_ipiv by the pivots, hB (ldb x nrhs) by the solution.
extern "C" int bmb200_dgbsv_host(bmb200_handle_t h, int64_t n, int64_t kl, int64_t ku, int64_t nrhs, double *hAB,
                                 int64_t ldab, int64_t *h_ipiv, double *hB, int64_t ldb, int *info)
{
    if (!h) return -1;
    if (n < 0) return -2;
    if (kl < 0) return -3;
    if (ku < 0) return -4;
    if (nrhs < 0) return -5;
    if (ldab < 2 * kl + ku + 1) return -7;
    if (ldb < imax64(1, n)) return -10;
    if (info) *info = 0;
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    const size_t nab = (size_t)ldab * n, nb = (size_t)ldb * (nrhs > 0 ? nrhs : 0);
    // own allocations (the factorisation may itself use the scratch arena)
    double *dAB = nullptr, *dB = nullptr;
    i64 *dip = nullptr;
    BMB_CUDA(h, cudaMalloc(&dAB, nab * sizeof(double)));
    BMB_CUDA(h, cudaMalloc(&dip, (size_t)n * sizeof(i64)));
    if (nb) BMB_CUDA(h, cudaMalloc(&dB, nb * sizeof(double)));
    cudaStream_t ks = h->stream, cs = h->copy_stream;
    int rc = 0, linfo = 0;
    do {
        if (cudaMemcpyAsync(dAB, hAB, nab * sizeof(double), cudaMemcpyHostToDevice, ks) != cudaSuccess) { rc = BMB200_ERR_CUDA; break; }
        if (nb) {  // B rides on the copy stream while the factorisation runs
            if (cudaMemcpyAsync(dB, hB, nb * sizeof(double), cudaMemcpyHostToDevice, cs) != cudaSuccess) { rc = BMB200_ERR_CUDA; break; }
            cudaEventRecord(h->ev[0], cs);
        }
        rc = bmb200_dgbtrf(h, n, n, kl, ku, dAB, ldab, dip, &linfo);
        if (rc) break;
        if (info) *info = linfo;
        if (cudaMemcpyAsync(hAB, dAB, nab * sizeof(double), cudaMemcpyDeviceToHost, cs) != cudaSuccess) { rc = BMB200_ERR_CUDA; break; }
        if (cudaMemcpyAsync(h_ipiv, dip, (size_t)n * sizeof(i64), cudaMemcpyDeviceToHost, cs) != cudaSuccess) { rc = BMB200_ERR_CUDA; break; }
        if (linfo == 0 && nb) {
            cudaStreamWaitEvent(ks, h->ev[0], 0);
            rc = bmb200_dgbtrs(h, 'N', n, kl, ku, nrhs, dAB, ldab, dip, dB, ldb);
            if (rc) break;
            if (cudaMemcpyAsync(hB, dB, nb * sizeof(double), cudaMemcpyDeviceToHost, ks) != cudaSuccess) { rc = BMB200_ERR_CUDA; break; }
        }
    } while (0);
    cudaStreamSynchronize(ks);
    cudaStreamSynchronize(cs);
    cudaFree(dAB);
    cudaFree(dip);
    if (dB) cudaFree(dB);
    if (rc == BMB200_ERR_CUDA) snprintf(h->err, sizeof(h->err), "dgbsv_host: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

// _gbmm! on host band arrays: A, B up, one launch, C back.
extern "C" int bmb200_dgbmm_bb_host(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au,
                                    int64_t Bl, int64_t Bu, int64_t Cl, int64_t Cu, double alpha, const double *hA,
                                    int64_t lda, const double *hB, int64_t ldb, double beta, double *hC, int64_t ldc)
{
    if (!h) return -1;
    if (n < 0) return -2;
    if (nu < 0) return -3;
    if (m < 0) return -4;
    if (lda < Al + Au + 1) return -13;
    if (ldb < Bl + Bu + 1) return -15;
    if (ldc < Cl + Cu + 1) return -18;
    if (n == 0 || m == 0) return 0;
    DeviceGuard g(h->device);
    const size_t na = (size_t)lda * nu, nb = (size_t)ldb * m, nc = (size_t)ldc * m;
    int rc = bmb_ensure_scratch(h, (na + nb + nc) * sizeof(double) + 3 * 512);
    if (rc) return rc;
    Arena ar(h->scratch);
    double *dA = ar.take<double>(na), *dB = ar.take<double>(nb), *dC = ar.take<double>(nc);
    cudaStream_t ks = h->stream, cs = h->copy_stream;
    BMB_CUDA(h, cudaEventRecord(h->ev[3], ks));
    BMB_CUDA(h, cudaStreamWaitEvent(cs, h->ev[3], 0));
    if (na) BMB_CUDA(h, cudaMemcpyAsync(dA, hA, na * sizeof(double), cudaMemcpyHostToDevice, ks));
    if (nb) BMB_CUDA(h, cudaMemcpyAsync(dB, hB, nb * sizeof(double), cudaMemcpyHostToDevice, cs));
    // C goes up when it is read (beta != 0) or when the written window is narrower than the column stride
    if (beta != 0.0 || ldc != Cl + Cu + 1) BMB_CUDA(h, cudaMemcpyAsync(dC, hC, nc * sizeof(double), cudaMemcpyHostToDevice, cs));
    BMB_CUDA(h, cudaEventRecord(h->ev[0], cs));
    BMB_CUDA(h, cudaStreamWaitEvent(ks, h->ev[0], 0));
    rc = bmb200_dgbmm_bb(h, n, nu, m, Al, Au, Bl, Bu, Cl, Cu, alpha, dA, lda, dB, ldb, beta, dC, ldc);
    if (rc) return rc;
    BMB_CUDA(h, cudaMemcpyAsync(hC, dC, nc * sizeof(double), cudaMemcpyDeviceToHost, ks));
    BMB_CUDA(h, cudaStreamSynchronize(ks));
    return 0;
}
