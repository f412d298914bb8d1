// context.cu -- handle, memory and small utility kernels of libbmb200.
#include "common.cuh"

extern "C" int bmb200_version(void) { return BMB200_VERSION; }

extern "C" int bmb200_create(bmb200_handle_t *out, int device, void *stream)
{
    if (!out) return -1;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        fprintf(stderr, "libbmb200: no CUDA device available -- there is no CPU fallback\n");
        return BMB200_ERR_CUDA - (int)cudaErrorNoDevice;
    }
    if (device < 0 || device >= count) return -2;
    bmb200_ctx *h = new bmb200_ctx();
    h->device = device;
    h->stream = (cudaStream_t)stream;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return BMB200_ERR_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        fprintf(stderr, "libbmb200: device %d is sm_%d%d; this library is built for sm_100a only\n", device,
                prop.major, prop.minor);
        delete h;
        return BMB200_ERR_CUDA - (int)cudaErrorInvalidDevice;
    }
    if (cudaMalloc(&h->d_info, 64 * sizeof(int)) != cudaSuccess) { delete h; return BMB200_ERR_CUDA; }
    cudaMemset(h->d_info, 0, 64 * sizeof(int));
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming);
    *out = h;
    return 0;
}

extern "C" int bmb200_destroy(bmb200_handle_t h)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    bmb200_halo_destroy(h);
    if (h->d_info) cudaFree(h->d_info);
    if (h->scratch) cudaFree(h->scratch);
    if (h->backup) cudaFree(h->backup);
    for (int i = 0; i < 2; ++i)
        if (h->pinned[i]) cudaFreeHost(h->pinned[i]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 4; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
    return 0;
}

extern "C" int bmb200_set_stream(bmb200_handle_t h, void *stream)
{
    if (!h) return -1;
    h->stream = (cudaStream_t)stream;
    return 0;
}

extern "C" int bmb200_sync(bmb200_handle_t h)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    // a sharded gbmv whose neighbour never published its halo raises a device-side flag instead of hanging
    int flag = 0;
    BMB_CUDA(h, cudaMemcpy(&flag, h->d_info + 8, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(h->d_info + 8, 0, sizeof(int));
        snprintf(h->err, sizeof(h->err), "dgbmv_sharded: timed out waiting for a neighbour's x halo");
        return BMB200_ERR_CUDA;
    }
    return 0;
}

extern "C" int bmb200_malloc(bmb200_handle_t h, void **dptr, size_t bytes)
{
    if (!h) return -1;
    if (!dptr) return -2;
    DeviceGuard g(h->device);
    BMB_CUDA(h, cudaMalloc(dptr, bytes ? bytes : 1));
    return 0;
}

extern "C" int bmb200_free(bmb200_handle_t h, void *dptr)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    BMB_CUDA(h, cudaFree(dptr));
    return 0;
}

extern "C" int bmb200_memcpy_h2d(bmb200_handle_t h, void *dst, const void *hsrc, size_t bytes)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    BMB_CUDA(h, cudaMemcpyAsync(dst, hsrc, bytes, cudaMemcpyHostToDevice, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bmb200_memcpy_d2h(bmb200_handle_t h, void *hdst, const void *dsrc, size_t bytes)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    BMB_CUDA(h, cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" const char *bmb200_last_error(bmb200_handle_t h) { return h ? h->err : "null handle"; }
extern "C" int64_t bmb200_launch_count(bmb200_handle_t h) { return h ? h->launches : -1; }

int bmb_ensure_scratch(bmb200_ctx *h, size_t bytes)
{
    if (bytes <= h->scratch_bytes) return 0;
    if (h->scratch) {
        BMB_CUDA(h, cudaStreamSynchronize(h->stream));
        BMB_CUDA(h, cudaFree(h->scratch));
        h->scratch = nullptr;
        h->scratch_bytes = 0;
    }
    BMB_CUDA(h, cudaMalloc(&h->scratch, bytes));
    h->scratch_bytes = bytes;
    return 0;
}

// ---- C <- beta*C (beta == 0 zero-fills): _fill_lmul! / _fill_rmul!, src/generic/utils.jl:29-31 ----
__global__ void fill_lmul_kernel(double beta, double *__restrict__ c, i64 rows, i64 cols, i64 ldc, i64 inc)
{
    i64 total = rows * cols;
    for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        i64 j = t / rows, i = t - j * rows;
        double *p = c + i * inc + j * ldc;
        *p = (beta == 0.0) ? 0.0 : __dmul_rn(beta, *p);
    }
}

extern "C" int bmb200_dfill_lmul(bmb200_handle_t h, double beta, double *dC, int64_t rows, int64_t cols,
                                 int64_t ldc, int64_t inc)
{
    if (!h) return -1;
    if (rows < 0) return -4;
    if (cols < 0) return -5;
    if (rows == 0 || cols == 0 || beta == 1.0) return 0;
    if (!dC) return -3;
    DeviceGuard g(h->device);
    i64 total = rows * cols;
    int blocks = (int)imin64(cdiv64(total, 256), (i64)h->sm_count * 16);
    fill_lmul_kernel<<<blocks, 256, 0, h->stream>>>(beta, dC, rows, cols, ldc, inc);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// ---- widening copy for lu(A): BandedMatrix{T}(A,(l,l+u)), src/banded/BandedLU.jl:110 --------------
__global__ void band_widen_kernel(i64 n, i64 l, i64 rows_src, const double *__restrict__ a, i64 lda,
                                  double *__restrict__ ab, i64 ldab)
{
    i64 rows_dst = rows_src + l;
    i64 total = rows_dst * n;
    for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        i64 j = t / rows_dst, r = t - j * rows_dst;
        ab[r + j * ldab] = (r < l) ? 0.0 : a[(r - l) + j * lda];
    }
}

extern "C" int bmb200_dband_widen(bmb200_handle_t h, int64_t n, int64_t l, int64_t u, const double *dA,
                                  int64_t lda, double *dAB, int64_t ldab)
{
    if (!h) return -1;
    if (n < 0) return -2;
    if (l < 0) return -3;
    if (l + u + 1 < 0) return -4;
    if (lda < l + u + 1) return -6;
    if (ldab < 2 * l + u + 1) return -8;
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    i64 total = (2 * l + u + 1) * n;
    int blocks = (int)imin64(cdiv64(total, 256), (i64)h->sm_count * 16);
    band_widen_kernel<<<blocks, 256, 0, h->stream>>>(n, l, l + u + 1, dA, lda, dAB, ldab);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// ---- lu(A) = lu!(BandedMatrix{T}(A,(l,l+u))), src/banded/BandedLU.jl:106-111: widening copy + factorisation in one call.
// Knowing the source lets the optimistic wide-band path (gbtrf_strip.cu) skip its device-side copy of the band: if an
// interchange turns out to be needed it simply widens again. ----
extern "C" int bmb200_dgbtrf_from(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, const double *dA, int64_t lda,
                                  double *dAB, int64_t ldab, int64_t *d_ipiv, int *info)
{
    if (!h) return -1;
    if (lda < kl + ku + 1) return -7;
    if (ldab < 2 * kl + ku + 1) return -9;
    if (info) *info = 0;
    if (m == 0 || n == 0) return 0;
    int rc = bmb200_dband_widen(h, n, kl, ku, dA, lda, dAB, ldab);
    if (rc) return rc;
    h->lu_src = dA;
    h->lu_src_ld = lda;
    rc = bmb200_dgbtrf(h, m, n, kl, ku, dAB, ldab, d_ipiv, info);
    h->lu_src = nullptr;
    return rc;
}

// ---- development hook (include/bmb200_internal.h): the only writer of the handle's tuning block ----
extern "C" int bmb200_internal_set_tuning(bmb200_handle_t h, const char *key, long long value)
{
    if (!h || !key) return -1;
    bmb_tuning &t = h->tune;
    struct { const char *name; int *slot; } ints[] = {
        {"gbmm_ring", &t.gbmm_ring}, {"gbmm_nt", &t.gbmm_nt}, {"gbmm_rw", &t.gbmm_rw}, {"gbtrf_nopipe", &t.gbtrf_nopipe}, {"gbtrf_nostrip", &t.gbtrf_nostrip}, {"gbtrf_nomw", &t.gbtrf_nomw},
        {"gbmm_wide", &t.gbmm_wide}, {"pb_nopdl", &t.pb_nopdl}, {"pb_nobulk", &t.pb_nobulk}, {"pb_clate", &t.pb_clate}, {"gbmv_spr", &t.gbmv_spr}, {"typed_nowin", &t.typed_nowin}, {"pb_nodiag", &t.pb_nodiag}, {"pipe_nospec", &t.pipe_nospec}, {"pipe_stats", &t.pipe_stats}, {"gbtrs_noblock", &t.gbtrs_noblock},
        {"gbtrs_pfdist_blocked", &t.gbtrs_pfdist_blocked}, {"gbtrs_nocluster", &t.gbtrs_nocluster}, {"gbtrs_nosplit", &t.gbtrs_nosplit},
        {"gbtrs_cluster", &t.gbtrs_cluster}, {"gbtrs_pfdist", &t.gbtrs_pfdist}, {"gbtrs_stats", &t.gbtrs_stats},
        {"debug", &t.debug}, {"sbmv_rows_k", &t.sbmv_rows_k},
    };
    for (auto &e : ints)
        if (strcmp(key, e.name) == 0) { *e.slot = (int)value; return 0; }
    if (strcmp(key, "pipe_maxpanels") == 0) { t.pipe_maxpanels = value; return 0; }
    if (strcmp(key, "reset") == 0) { t = bmb_tuning(); return 0; }
    return -2;
}
