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
    if (m == 0 || n == 0 || xl + xu + 1 <= 0 || yl + yu + 1 <= 0) return 0;   // no bands in X or Y (broadcast.jl:987,1008)
    if (!dX || !dY) return -6;
    DeviceGuard g(h->device);
    const BandDesc X{m, n, (int)xl, (int)xu, ldx}, Y{m, n, (int)yl, (int)yu, ldy};
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
    if (m == 0 || n == 0 || dl + du + 1 <= 0) return 0;
    if (!dD || (sl + su + 1 > 0 && !dS)) return -6;
    DeviceGuard g(h->device);
    const BandDesc S{m, n, (int)sl, (int)su, lds}, D{m, n, (int)dl, (int)du, ldd};
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
