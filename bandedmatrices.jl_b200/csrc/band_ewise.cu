// band_ewise.cu -- band-aligned elementwise operations between BandedMatrices of DIFFERENT bandwidths, on the device: the steps
// either side of the hot path that build e.g. I - dt*Laplacian (examples/finitedifference_2d.jl:15,29)
// and today are scalar host loops in the reference:
//   * bmb200_dband_axpy  = banded_axpy!(a, X, Y) (src/banded/BandedMatrix.jl:1006-1015, src/generic/broadcast.jl:978-1020):
//       equal bandwidths  -> axpy!(a, X.data, Y.data): ONE fused multiply-add per slot of the data arrays (OpenBLAS daxpy),
//                            corner slots included;
//       otherwise         -> Y[k,j] = a*X[k,j] + Y[k,j] (a rounded product, then a sum: the reference's scalar loop) on the
//                            overlapping bands; BandError if X has a non-zero entry in a band Y does not store.
//   * bmb200_dband_copy  = copyto!(dest, src) between bandwidths (_banded_broadcast!(dest, identity, src), broadcast.jl:175-230):
//       overlapping bands copied, dest's other in-matrix band entries zeroed, BandError for non-zeros of src outside dest's bands;
//       corner slots of dest are left alone.
// Both are pure HBM streaming (24 / 16 bytes per stored entry).  A 2-D thread block maps (band row, column) so that a warp
// covers 32/RX adjacent columns of RX band rows: contiguous memory whenever ld equals the number of band rows.
#include "common.cuh"

struct BandDesc {
    i64 m, n;      // matrix size
    int l, u;      // bandwidths
    i64 ld;        // leading dimension of the data array
};

// non-zero entries of X (in-matrix) that lie in bands the other matrix does not store
__global__ void __launch_bounds__(256)
band_count_outside(BandDesc X, const double *__restrict__ x, int yl, int yu, int RX, unsigned long long *__restrict__ cnt)
{
    const int tx = threadIdx.x % RX, ty = threadIdx.x / RX, CY = blockDim.x / RX;
    const int R = X.l + X.u + 1;
    unsigned long long c = 0;
    for (i64 j = (i64)blockIdx.x * CY + ty; j < X.n; j += (i64)gridDim.x * CY)
        for (int r = tx; r < R; r += RX) {
            const i64 k = j - X.u + r;            // matrix row
            const int d = (int)(k - j);           // k - j in [-u, l]
            if (k >= 0 && k < X.m && (d > yl || -d > yu) && x[r + j * X.ld] != 0.0) ++c;
        }
    if (c) atomicAdd(cnt, c);
}

// MODE 0: Y = fma(a, X, Y) on every slot (equal bandwidths); MODE 1: Y = a*X + Y on overlapping in-matrix entries;
// MODE 2: Y = X on overlapping in-matrix entries, 0 on Y's other in-matrix entries
template <int MODE>
__global__ void __launch_bounds__(256)
band_ewise(BandDesc X, const double *__restrict__ x, BandDesc Y, double *__restrict__ y, double a, int RX)
{
    const int tx = threadIdx.x % RX, ty = threadIdx.x / RX, CY = blockDim.x / RX;
    const int RY = Y.l + Y.u + 1, sh = X.u - Y.u;  // Y band row r <-> X band row r + sh
    const int RXrows = X.l + X.u + 1;
    for (i64 j = (i64)blockIdx.x * CY + ty; j < Y.n; j += (i64)gridDim.x * CY)
        for (int r = tx; r < RY; r += RX) {
            double *py = y + r + j * Y.ld;
            if (MODE == 0) {
                *py = fma(a, x[r + j * X.ld], *py);
                continue;
            }
            const i64 k = j - Y.u + r;
            if (k < 0 || k >= Y.m) continue;      // corner slot of Y: untouched
            const int rx = r + sh;
            const bool inx = rx >= 0 && rx < RXrows;
            if (MODE == 1) {
                if (inx) *py = __dadd_rn(__dmul_rn(a, x[rx + j * X.ld]), *py);
            } else {
                *py = inx ? x[rx + j * X.ld] : 0.0;
            }
        }
}

static int rx_for(int rows) { int r = 1; while (r < rows && r < 32) r <<= 1; return r; }

static int band_check(bmb200_ctx *h, const BandDesc &X, const double *dX, int yl, int yu, int64_t *nonzero_outside)
{
    *nonzero_outside = 0;
    if (X.l <= yl && X.u <= yu) return 0;
    unsigned long long *cnt = (unsigned long long *)(h->d_info + 20);
    BMB_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), h->stream));
    const int RX = rx_for(X.l + X.u + 1);
    const i64 blocks = imin64(cdiv64(X.n, 256 / RX), (i64)h->sm_count * 16);
    band_count_outside<<<(unsigned)blocks, 256, 0, h->stream>>>(X, dX, yl, yu, RX, cnt);
    BMB_LAUNCH_CHECK(h);
    unsigned long long hc = 0;
    BMB_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    *nonzero_outside = (int64_t)hc;
    return 0;
}

static int band_args(int64_t m, int64_t n, int64_t xl, int64_t xu, int64_t ldx, int64_t yl, int64_t yu, int64_t ldy)
{
    if (m < 0) return -2;
    if (n < 0) return -3;
    if (xl + xu + 1 < 0 || xl + xu + 1 >= ((int64_t)1 << 30)) return -4;
    if (yl + yu + 1 < 0 || yl + yu + 1 >= ((int64_t)1 << 30)) return -8;
    if (ldx < imax64(1, xl + xu + 1)) return -7;
    if (ldy < imax64(1, yl + yu + 1)) return -11;
    return 0;
}

// *nonzero_outside (host) receives the number of non-zero entries of X in bands Y does not store; when it is not 0 nothing is
// written (the reference throws BandError before touching Y) and the call returns 0.
extern "C" int bmb200_dband_axpy(bmb200_handle_t h, int64_t m, int64_t n, double a, int64_t xl, int64_t xu, const double *dX, int64_t ldx,
                                 int64_t yl, int64_t yu, double *dY, int64_t ldy, int64_t *nonzero_outside)
{
    if (!h) return -1;
    const int rc0 = band_args(m, n, xl, xu, ldx, yl, yu, ldy);
    if (rc0) return rc0;
    if (!nonzero_outside) return -13;
    *nonzero_outside = 0;
    if (m == 0 || n == 0 || xl + xu + 1 <= 0) return 0;   // no bands in X (broadcast.jl:987)
    if (!dX) return -6;
    DeviceGuard g(h->device);
    const BandDesc X{m, n, (int)xl, (int)xu, ldx}, Y{m, n, (int)yl, (int)yu, ldy};
    if (yl + yu + 1 <= 0) {  // Y stores no band: the reference still looks at X first (broadcast.jl:990-1006), then returns (:1008)
        return band_check(h, X, dX, (int)yl, (int)yu, nonzero_outside);
    }
    if (!dY) return -6;
    const int RX = rx_for(Y.l + Y.u + 1);
    const i64 blocks = imin64(cdiv64(n, 256 / RX), (i64)h->sm_count * 16);
    if (xl == yl && xu == yu) {
        band_ewise<0><<<(unsigned)blocks, 256, 0, h->stream>>>(X, dX, Y, dY, a, RX);
        BMB_LAUNCH_CHECK(h);
        return 0;
    }
    const int rc = band_check(h, X, dX, (int)yl, (int)yu, nonzero_outside);
    if (rc || *nonzero_outside) return rc;
    band_ewise<1><<<(unsigned)blocks, 256, 0, h->stream>>>(X, dX, Y, dY, a, RX);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int bmb200_dband_copy(bmb200_handle_t h, int64_t m, int64_t n, int64_t sl, int64_t su, const double *dS, int64_t lds, int64_t dl,
                                 int64_t du, double *dD, int64_t ldd, int64_t *nonzero_outside)
{
    if (!h) return -1;
    const int rc0 = band_args(m, n, sl, su, lds, dl, du, ldd);
    if (rc0) return rc0;
    if (!nonzero_outside) return -12;
    *nonzero_outside = 0;
    if (m == 0 || n == 0) return 0;
    if ((dl + du + 1 > 0 && !dD) || (sl + su + 1 > 0 && !dS)) return -6;
    DeviceGuard g(h->device);
    const BandDesc S{m, n, (int)sl, (int)su, lds}, D{m, n, (int)dl, (int)du, ldd};
    if (dl + du + 1 <= 0) {  // an empty-band destination still raises BandError for a non-zero source (same order as the reference)
        return (sl + su + 1 > 0) ? band_check(h, S, dS, (int)dl, (int)du, nonzero_outside) : 0;
    }
    if (sl + su + 1 > 0) {
        const int rc = band_check(h, S, dS, (int)dl, (int)du, nonzero_outside);
        if (rc || *nonzero_outside) return rc;
    }
    const int RX = rx_for(D.l + D.u + 1);
    const i64 blocks = imin64(cdiv64(n, 256 / RX), (i64)h->sm_count * 16);
    band_ewise<2><<<(unsigned)blocks, 256, 0, h->stream>>>(S, dS, D, dD, 0.0, RX);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Band utilities the gbmm! driver and the broadcasting layer need on the device (no host round trips, no eager tensor ops):
//   band_lmul_block   lmul!(beta, view(C, r0+1:r1, c0+1:c1)) on the stored band (gbmm.jl:234-249: the rows / columns a
//                     negative-bandwidth operand leaves untouched); beta == 0 zero-fills
//   band_transpose    convert(BandedMatrix, A') (matmul.jl:182-184): band row r of A' is band row l+u-r of A, shifted
//   band_nonzero_rows one flag per band row: does it hold a non-zero in-matrix entry (gbmm.jl:191-205 counts the leading /
//                     trailing all-zero bands from these)
//   band_axpby        Z = alpha*X + beta*Y entry by entry over Z's band (X, Y read as 0 outside their own bands): the
//                     arithmetic of the reference's broadcast kernels for A .+ B, A .- B, a .* A .+ b .* B
//                     (src/generic/broadcast.jl:359-384, 927-964): products and the sum are rounded separately
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
band_lmul_block(BandDesc Cd, double *__restrict__ c, i64 r0, i64 r1, i64 c0, i64 c1, double beta, int RX)
{
    const int tx = threadIdx.x % RX, ty = threadIdx.x / RX, CY = blockDim.x / RX;
    const int R = Cd.l + Cd.u + 1;
    for (i64 j = c0 + (i64)blockIdx.x * CY + ty; j < c1; j += (i64)gridDim.x * CY)
        for (int r = tx; r < R; r += RX) {
            const i64 k = j - Cd.u + r;
            if (k < r0 || k >= r1) continue;
            double *p = c + r + j * Cd.ld;
            *p = (beta == 0.0) ? 0.0 : __dmul_rn(beta, *p);
        }
}

__global__ void __launch_bounds__(256)
band_transpose(BandDesc S, const double *__restrict__ s, BandDesc D, double *__restrict__ d, int RX)
{
    // D = S' : D is S.n x S.m with (S.u, S.l); D[k', j'] = S[j', k']  ->  d[(D.u + k' - j') + j'*D.ld] = s[(S.u + j' - k') + k'*S.ld]
    const int tx = threadIdx.x % RX, ty = threadIdx.x / RX, CY = blockDim.x / RX;
    const int R = D.l + D.u + 1;
    for (i64 j = (i64)blockIdx.x * CY + ty; j < D.n; j += (i64)gridDim.x * CY)
        for (int r = tx; r < R; r += RX) {
            const i64 k = j - D.u + r;  // row of D = column of S
            d[r + j * D.ld] = (k >= 0 && k < D.m) ? s[(S.u + j - k) + k * S.ld] : 0.0;
        }
}

__global__ void __launch_bounds__(256)
band_nonzero_rows(BandDesc X, const double *__restrict__ x, int RX, int *__restrict__ flags)
{
    const int tx = threadIdx.x % RX, ty = threadIdx.x / RX, CY = blockDim.x / RX;
    const int R = X.l + X.u + 1;
    for (int r = tx; r < R; r += RX) {
        bool nz = false;
        for (i64 j = (i64)blockIdx.x * CY + ty; j < X.n && !nz; j += (i64)gridDim.x * CY) {
            const i64 k = j - X.u + r;
            nz = k >= 0 && k < X.m && x[r + j * X.ld] != 0.0;
        }
        if (nz) flags[r] = 1;
    }
}

__global__ void __launch_bounds__(256)
band_axpby(BandDesc X, const double *__restrict__ x, BandDesc Y, const double *__restrict__ y, BandDesc Z, double *__restrict__ z,
           double alpha, double beta, int RX)
{
    const int tx = threadIdx.x % RX, ty = threadIdx.x / RX, CY = blockDim.x / RX;
    const int R = Z.l + Z.u + 1;
    for (i64 j = (i64)blockIdx.x * CY + ty; j < Z.n; j += (i64)gridDim.x * CY)
        for (int r = tx; r < R; r += RX) {
            const i64 k = j - Z.u + r;
            if (k < 0 || k >= Z.m) continue;  // corner slot: untouched
            const int dgl = (int)(k - j);
            const bool inx = dgl <= X.l && -dgl <= X.u, iny = dgl <= Y.l && -dgl <= Y.u;
            const double xv = inx ? x[(X.u + dgl) + j * X.ld] : 0.0;
            const double yv = iny ? y[(Y.u + dgl) + j * Y.ld] : 0.0;
            double v;
            if (beta == 0.0) v = __dmul_rn(alpha, xv);
            else if (alpha == 1.0 && beta == 1.0) v = __dadd_rn(xv, yv);
            else v = __dadd_rn(__dmul_rn(alpha, xv), __dmul_rn(beta, yv));
            z[r + j * Z.ld] = v;
        }
}

static int band_rx(int rows) { return rows >= 32 ? 32 : rows >= 16 ? 16 : rows >= 8 ? 8 : rows >= 4 ? 4 : rows >= 2 ? 2 : 1; }
static unsigned band_grid(bmb200_ctx *h, i64 cols, int rx) { return (unsigned)imin64(cdiv64(cols, 256 / rx), (i64)h->sm_count * 16); }

extern "C" int bmb200_dband_lmul_block(bmb200_handle_t h, int64_t m, int64_t n, int64_t l, int64_t u, double *dC, int64_t ldc,
                                       int64_t r0, int64_t r1, int64_t c0, int64_t c1, double beta)
{
    if (!h) return -1;
    if (m < 0 || n < 0) return -2;
    const i64 rows = l + u + 1;
    if (rows <= 0 || r1 <= r0 || c1 <= c0 || m == 0 || n == 0) return 0;
    if (ldc < rows) return -7;
    if (r0 < 0) r0 = 0;
    if (r1 > m) r1 = m;
    if (c0 < 0) c0 = 0;
    if (c1 > n) c1 = n;
    if (r1 <= r0 || c1 <= c0) return 0;
    DeviceGuard g(h->device);
    const int rx = band_rx((int)rows);
    band_lmul_block<<<band_grid(h, c1 - c0, rx), 256, 0, h->stream>>>(BandDesc{m, n, (int)l, (int)u, ldc}, dC, r0, r1, c0, c1, beta, rx);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int bmb200_dband_transpose(bmb200_handle_t h, int64_t m, int64_t n, int64_t l, int64_t u, const double *dS, int64_t lds,
                                      double *dD, int64_t ldd)
{
    if (!h) return -1;
    if (m < 0 || n < 0) return -2;
    const i64 rows = l + u + 1;
    if (rows <= 0 || m == 0 || n == 0) return 0;
    if (lds < rows) return -7;
    if (ldd < rows) return -9;
    DeviceGuard g(h->device);
    const int rx = band_rx((int)rows);
    band_transpose<<<band_grid(h, m, rx), 256, 0, h->stream>>>(BandDesc{m, n, (int)l, (int)u, lds}, dS, BandDesc{n, m, (int)u, (int)l, ldd}, dD, rx);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// flags_host[r] (r = 0 .. l+u) = 1 when band row r of the data array holds a non-zero in-matrix entry; synchronises.
extern "C" int bmb200_dband_nonzero_rows(bmb200_handle_t h, int64_t m, int64_t n, int64_t l, int64_t u, const double *dX, int64_t ldx,
                                         int *flags_host)
{
    if (!h) return -1;
    const i64 rows = l + u + 1;
    if (rows <= 0) return 0;
    if (rows > 4096) return -4;
    if (ldx < rows) return -7;
    DeviceGuard g(h->device);
    if (int rc = bmb_ensure_scratch(h, (size_t)rows * sizeof(int))) return rc;
    int *d_flags = static_cast<int *>(h->scratch);
    BMB_CUDA(h, cudaMemsetAsync(d_flags, 0, (size_t)rows * sizeof(int), h->stream));
    if (m > 0 && n > 0) {
        const int rx = band_rx((int)rows);
        band_nonzero_rows<<<band_grid(h, n, rx), 256, 0, h->stream>>>(BandDesc{m, n, (int)l, (int)u, ldx}, dX, rx, d_flags);
        BMB_LAUNCH_CHECK(h);
    }
    BMB_CUDA(h, cudaMemcpyAsync(flags_host, d_flags, (size_t)rows * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int bmb200_dband_axpby(bmb200_handle_t h, int64_t m, int64_t n, double alpha, int64_t xl, int64_t xu, const double *dX,
                                  int64_t ldx, double beta, int64_t yl, int64_t yu, const double *dY, int64_t ldy, int64_t zl,
                                  int64_t zu, double *dZ, int64_t ldz)
{
    if (!h) return -1;
    if (m < 0 || n < 0) return -2;
    const i64 rows = zl + zu + 1;
    if (rows <= 0 || m == 0 || n == 0) return 0;
    if (ldz < rows) return -17;
    if (xl + xu + 1 > 0 && ldx < xl + xu + 1) return -8;
    if (yl + yu + 1 > 0 && ldy < yl + yu + 1) return -13;
    DeviceGuard g(h->device);
    const int rx = band_rx((int)rows);
    band_axpby<<<band_grid(h, n, rx), 256, 0, h->stream>>>(BandDesc{m, n, (int)xl, (int)xu, ldx}, dX, BandDesc{m, n, (int)yl, (int)yu, ldy}, dY,
                                                        BandDesc{m, n, (int)zl, (int)zu, ldz}, dZ, alpha, beta, rx);
    BMB_LAUNCH_CHECK(h);
    return 0;
}
