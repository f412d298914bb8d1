// typed.cu -- the S / C / Z instantiations of the band BLAS / LAPACK entry points (src/blas.jl:4-7 generates gbmv! / sbmv! / hbmv!
// for Float32, ComplexF32 and ComplexF64 besides Float64; LAPACK.gbtrf! / gbtrs! take the same four element types,
// src/banded/BandedLU.jl:90-103, src/banded/linalg.jl:24-63 with a true conjugate-transpose solve at :57-63).  SURVEY.md 8(f)
// rank 1.  The Float64 path keeps its tuned kernels (gbmv.cu, gbtrf_*.cu, gbtrs_*.cu); here ONE generic kernel per operation
// is instantiated for float, complex<float> and complex<double> -- and for double as the check against the tuned path:
//   tgbmv_n      y <- alpha*A*x + beta*y, one thread per row, columns ascending, t = alpha*x[j] rounded first (reference order)
//   tgbmv_t      y <- alpha*op(A)*x + beta*y, op = transpose or conjugate transpose, one warp per column of A
//   thbmv        Hermitian (real types: symmetric) band matvec from one stored triangle, one thread per row
//   tgbtf2       unblocked partial-pivot band LU in LAPACK's xGBTF2 order (pivot = first maximum of |re|+|im|, multipliers scaled
//                by the reciprocal of the pivot, rank-1 update), one CTA, three barriers per column
//   tgbtrs       solve with the factors for op in {N, T, C}, one CTA per right-hand side
//   ttbmv/ttbsv  triangular band multiply / solve for op in {N, T, C} (tbmv! / tbsv!, src/blas.jl:71-141)
//   tpbtf2       Hermitian band Cholesky in xPBTF2's order, one CTA; pbtrs = two ttbsv sweeps (pbtrf! / pbtrs!, src/lapack.jl:268-332)
//   tgbmm_bb/bd  banded x banded and banded x dense, one thread per entry of C, inner index ascending (gbmm!, src/banded/gbmm.jl)
// OpenBLAS' operation order is unspecified for these types (SIMD dot / complex kernels): parity is to rounding (tests: 1e-5 /
// 1e-13 relative for single / double precision), pivots compared exactly on well-separated columns.
#include <cuComplex.h>

#include "common.cuh"

namespace {

template <typename T> struct Num;
template <> struct Num<float> {
    typedef float real;
    static __device__ __forceinline__ float zero() { return 0.0f; }
    static __device__ __forceinline__ float one() { return 1.0f; }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float neg(float a) { return -a; }
    static __device__ __forceinline__ float conj(float a) { return a; }
    static __device__ __forceinline__ float abs1(float a) { return fabsf(a); }
    static __device__ __forceinline__ float recip(float a) { return 1.0f / a; }
    static __device__ __forceinline__ bool iszero(float a) { return a == 0.0f; }
    static __device__ __forceinline__ float realpart(float a) { return a; }
    static __device__ __forceinline__ float shfl_xor(float a, int o) { return __shfl_xor_sync(0xffffffffu, a, o); }
};
template <> struct Num<double> {
    typedef double real;
    static __device__ __forceinline__ double zero() { return 0.0; }
    static __device__ __forceinline__ double one() { return 1.0; }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double neg(double a) { return -a; }
    static __device__ __forceinline__ double conj(double a) { return a; }
    static __device__ __forceinline__ double abs1(double a) { return fabs(a); }
    static __device__ __forceinline__ double recip(double a) { return 1.0 / a; }
    static __device__ __forceinline__ bool iszero(double a) { return a == 0.0; }
    static __device__ __forceinline__ double realpart(double a) { return a; }
    static __device__ __forceinline__ double shfl_xor(double a, int o) { return __shfl_xor_sync(0xffffffffu, a, o); }
};
template <> struct Num<cuFloatComplex> {
    typedef float real;
    typedef cuFloatComplex T;
    static __device__ __forceinline__ T zero() { return make_cuFloatComplex(0.0f, 0.0f); }
    static __device__ __forceinline__ T one() { return make_cuFloatComplex(1.0f, 0.0f); }
    static __device__ __forceinline__ T mul(T a, T b) { return make_cuFloatComplex(fmaf(a.x, b.x, -__fmul_rn(a.y, b.y)), fmaf(a.x, b.y, __fmul_rn(a.y, b.x))); }
    static __device__ __forceinline__ T fma(T a, T b, T c) { return make_cuFloatComplex(fmaf(-a.y, b.y, fmaf(a.x, b.x, c.x)), fmaf(a.y, b.x, fmaf(a.x, b.y, c.y))); }
    static __device__ __forceinline__ T add(T a, T b) { return make_cuFloatComplex(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
    static __device__ __forceinline__ T neg(T a) { return make_cuFloatComplex(-a.x, -a.y); }
    static __device__ __forceinline__ T conj(T a) { return make_cuFloatComplex(a.x, -a.y); }
    static __device__ __forceinline__ float abs1(T a) { return fabsf(a.x) + fabsf(a.y); }
    static __device__ __forceinline__ T recip(T a) { const float d = fmaf(a.x, a.x, __fmul_rn(a.y, a.y)); return make_cuFloatComplex(a.x / d, -a.y / d); }
    static __device__ __forceinline__ bool iszero(T a) { return a.x == 0.0f && a.y == 0.0f; }
    static __device__ __forceinline__ T realpart(T a) { return make_cuFloatComplex(a.x, 0.0f); }
    static __device__ __forceinline__ T shfl_xor(T a, int o) { return make_cuFloatComplex(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o)); }
};
template <> struct Num<cuDoubleComplex> {
    typedef double real;
    typedef cuDoubleComplex T;
    static __device__ __forceinline__ T zero() { return make_cuDoubleComplex(0.0, 0.0); }
    static __device__ __forceinline__ T one() { return make_cuDoubleComplex(1.0, 0.0); }
    static __device__ __forceinline__ T mul(T a, T b) { return make_cuDoubleComplex(::fma(a.x, b.x, -__dmul_rn(a.y, b.y)), ::fma(a.x, b.y, __dmul_rn(a.y, b.x))); }
    static __device__ __forceinline__ T fma(T a, T b, T c) { return make_cuDoubleComplex(::fma(-a.y, b.y, ::fma(a.x, b.x, c.x)), ::fma(a.y, b.x, ::fma(a.x, b.y, c.y))); }
    static __device__ __forceinline__ T add(T a, T b) { return make_cuDoubleComplex(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
    static __device__ __forceinline__ T neg(T a) { return make_cuDoubleComplex(-a.x, -a.y); }
    static __device__ __forceinline__ T conj(T a) { return make_cuDoubleComplex(a.x, -a.y); }
    static __device__ __forceinline__ double abs1(T a) { return fabs(a.x) + fabs(a.y); }
    static __device__ __forceinline__ T recip(T a) { const double d = ::fma(a.x, a.x, __dmul_rn(a.y, a.y)); return make_cuDoubleComplex(a.x / d, -a.y / d); }
    static __device__ __forceinline__ bool iszero(T a) { return a.x == 0.0 && a.y == 0.0; }
    static __device__ __forceinline__ T realpart(T a) { return make_cuDoubleComplex(a.x, 0.0); }
    static __device__ __forceinline__ T shfl_xor(T a, int o) { return make_cuDoubleComplex(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o)); }
};

// ---- gbmv ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
tgbmv_n(i64 m, i64 n, i64 kl, i64 ku, T alpha, const T *__restrict__ a, i64 lda, const T *__restrict__ x, i64 incx, T beta, T *__restrict__ y, i64 incy)
{
    typedef Num<T> N;
    const T *x0 = incx < 0 ? x - (n - 1) * incx : x;
    T *y0 = incy < 0 ? y - (m - 1) * incy : y;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        T acc = N::iszero(beta) ? N::zero() : N::mul(beta, y0[i * incy]);
        if (!N::iszero(alpha)) {
            const i64 jlo = i - kl > 0 ? i - kl : 0, jhi = i + ku < n - 1 ? i + ku : n - 1;
            const T *p = a + (ku + i - jlo) + jlo * lda;  // A[i,j]; next column: + lda - 1
            for (i64 j = jlo; j <= jhi; ++j, p += lda - 1) acc = N::fma(N::mul(alpha, x0[j * incx]), *p, acc);
        }
        y0[i * incy] = acc;
    }
}

template <typename T, bool CONJ>
__global__ void __launch_bounds__(256)
tgbmv_t(i64 m, i64 n, i64 kl, i64 ku, T alpha, const T *__restrict__ a, i64 lda, const T *__restrict__ x, i64 incx, T beta, T *__restrict__ y, i64 incy)
{
    typedef Num<T> N;
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const T *x0 = incx < 0 ? x - (m - 1) * incx : x;
    T *y0 = incy < 0 ? y - (n - 1) * incy : y;
    for (i64 j = warp; j < n; j += nwarps) {
        const i64 ilo = j - ku > 0 ? j - ku : 0, ihi = j + kl < m - 1 ? j + kl : m - 1;
        const T *col = a + j * lda + (ku - j);  // A[i,j] = col[i]
        T acc = N::zero();
        if (!N::iszero(alpha))
            for (i64 i = ilo + lane; i <= ihi; i += 32) acc = N::fma(CONJ ? N::conj(col[i]) : col[i], x0[i * incx], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = N::add(acc, N::shfl_xor(acc, o));
        if (lane == 0) {
            const T yb = N::iszero(beta) ? N::zero() : N::mul(beta, y0[j * incy]);
            y0[j * incy] = N::fma(alpha, acc, yb);
        }
    }
}

// narrow bands: one THREAD per column of A -- a column is a few contiguous elements (one or two 32-byte sectors, read whole by
// its thread), where the warp-per-column kernel would leave most lanes idle (measured at (4,3): 0.24-0.72 TB/s)
template <typename T, bool CONJ>
__global__ void __launch_bounds__(256)
tgbmv_t_thread(i64 m, i64 n, i64 kl, i64 ku, T alpha, const T *__restrict__ a, i64 lda, const T *__restrict__ x, i64 incx, T beta, T *__restrict__ y, i64 incy)
{
    typedef Num<T> N;
    const T *x0 = incx < 0 ? x - (m - 1) * incx : x;
    T *y0 = incy < 0 ? y - (n - 1) * incy : y;
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) {
        const i64 ilo = j - ku > 0 ? j - ku : 0, ihi = j + kl < m - 1 ? j + kl : m - 1;
        const T *col = a + j * lda + (ku - j);  // A[i,j] = col[i]
        T acc = N::zero();
        if (!N::iszero(alpha))
            for (i64 i = ilo; i <= ihi; ++i) acc = N::fma(CONJ ? N::conj(col[i]) : col[i], x0[i * incx], acc);
        const T yb = N::iszero(beta) ? N::zero() : N::mul(beta, y0[j * incy]);
        y0[j * incy] = N::fma(alpha, acc, yb);
    }
}

// ---- hbmv / sbmv: y <- alpha*H*x + beta*y, H Hermitian with one triangle in triangular-band storage ------------------------
template <typename T>
__global__ void __launch_bounds__(256)
thbmv(int up, i64 n, i64 k, T alpha, const T *__restrict__ a, i64 lda, const T *__restrict__ x, T beta, T *__restrict__ y)
{
    typedef Num<T> N;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const T yb = N::iszero(beta) ? N::zero() : N::mul(beta, y[i]);
        if (N::iszero(alpha)) { y[i] = yb; continue; }
        const T *col = a + i * lda + (up ? k : 0);  // diagonal entry of the own column
        T acc = N::mul(N::realpart(col[0]), x[i]);  // xHBMV reads only the real part of the diagonal
        for (i64 d = 1; d <= k; ++d) {
            const i64 js = up ? i - d : i + d;       // stored side: own column, H[js,i] stored => H[i,js] = conj
            const i64 jo = up ? i + d : i - d;       // other side: H[i,jo] stored in column jo
            if (js >= 0 && js < n) acc = N::fma(N::conj(up ? col[-d] : col[d]), x[js], acc);
            if (jo >= 0 && jo < n) acc = N::fma(a[(up ? k - d : d) + jo * lda], x[jo], acc);
        }
        y[i] = N::fma(alpha, acc, yb);
    }
}

// ---- gbtf2: unblocked band LU, LAPACK order.  ab is LU storage (ldab >= 2kl+ku+1), kv = kl+ku ------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
tgbtf2(i64 m, i64 n, i64 kl, i64 ku, T *__restrict__ ab, i64 ldab, i64 *__restrict__ ipiv, int *__restrict__ d_info)
{
    typedef Num<T> N;
    typedef typename N::real R;
    __shared__ R s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_jp;
    __shared__ T s_piv;
    const i64 kv = kl + ku, mn = m < n ? m : n;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    // fill-in rows of columns ku+1 .. min(kv, n)-1 start as zeros (DGBTF2 lines "Set fill-in elements ... to zero")
    for (i64 c = ku + 1; c < (kv < n ? kv : n); ++c)
        for (i64 r = kv - c + tid; r < kl; r += nt) ab[r + c * ldab] = N::zero();
    __syncthreads();
    i64 ju = 0;
    int info = 0;
    for (i64 j = 0; j < mn; ++j) {
        if (j + kv < n)
            for (i64 r = tid; r < kl; r += nt) ab[r + (j + kv) * ldab] = N::zero();
        const i64 km = kl < m - 1 - j ? kl : m - 1 - j;
        T *colj = ab + kv + j * ldab;  // colj[t] = A[j+t, j]
        // pivot: first maximum of |re|+|im| over t = 0..km
        R best = (R)-1;
        int bidx = 0;
        for (i64 t = tid; t <= km; t += nt) {
            const R v = N::abs1(colj[t]);
            if (v > best) { best = v; bidx = (int)t; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const R ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bidx; }
        __syncthreads();
        if (wid == 0) {
            best = lane < (nt >> 5) ? s_val[lane] : (R)-1;
            bidx = lane < (nt >> 5) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const R ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
            }
            if (lane == 0) {
                if (!(best >= (R)0) || bidx > km) bidx = 0;  // all-NaN column: LAPACK's IxAMAX returns the first index
                s_jp = bidx;
                s_piv = colj[bidx];
                ipiv[j] = j + bidx + 1;
            }
        }
        __syncthreads();
        const int jp = s_jp;
        const T piv = s_piv;
        if (!N::iszero(piv)) {
            const i64 jun = j + ku + jp < n - 1 ? j + ku + jp : n - 1;
            if (jun > ju) ju = jun;
            // row interchange over columns j..ju, then the multipliers
            if (jp != 0)
                for (i64 c = j + tid; c <= ju; c += nt) {
                    T *e = ab + kv - (c - j) + c * ldab;  // A[j, c]; A[j+jp, c] = e[jp]
                    const T t0 = e[0];
                    e[0] = e[jp];
                    e[jp] = t0;
                }
            __syncthreads();
            const T rinv = N::recip(piv);
            for (i64 t = 1 + tid; t <= km; t += nt) colj[t] = N::mul(colj[t], rinv);
            __syncthreads();
            // rank-1 update of the trailing block: A[j+t, c] -= l_t * A[j, c]
            const i64 ncol = ju - j;
            if (km > 0 && ncol > 0) {
                const i64 total = km * ncol;
                for (i64 e = tid; e < total; e += nt) {
                    const i64 cc = e / km, t = e - cc * km + 1, c = j + 1 + cc;
                    T *pc = ab + kv - (c - j) + c * ldab;  // A[j, c]
                    pc[t] = N::fma(N::neg(colj[t]), pc[0], pc[t]);
                }
            }
            __syncthreads();
        } else if (info == 0) {
            info = (int)(j + 1);
        }
    }
    if (tid == 0) d_info[0] = info;
}

// ---- gbtf2 for narrow bands: the kv+1 live columns of the LU storage sit in a shared-memory ring (whole columns of
// RP = 2kl+ku+1 entries, fill-in rows zeroed on entry as xGBTF2 does), so a step touches global memory only to retire column j and
// to fetch column j+kv+PF.  Three barriers per step: pivot search (warp 0) | every thread computes the NEW value of its entries
// of the (km+1) x (ju-j+1) block from OLD values only (row interchange, reciprocal scaling and rank-1 update fused: l_t and u_c are
// recomputed per entry) | write.  Same operation order per entry as the global-memory kernel above.
#define TW_PF 4
template <typename T, int NT, int E>
__global__ void __launch_bounds__(NT)
tgbtf2_win(i64 m, i64 n, int kl, int ku, T *__restrict__ ab, i64 ldab, i64 *__restrict__ ipiv, int *__restrict__ d_info, int slots, int RP)
{
    typedef Num<T> N;
    typedef typename N::real R;
    extern __shared__ __align__(16) unsigned char tw_raw[];
    T *win = reinterpret_cast<T *>(tw_raw);
    __shared__ int s_jp;
    __shared__ T s_piv;
    const int kv = kl + ku, tid = threadIdx.x, lane = tid & 31, SM = slots - 1;
    const i64 mn = m < n ? m : n;
    auto colp = [&](i64 c) { return win + (size_t)((int)c & SM) * RP; };
    // column c of the LU storage -> its slot, ASYNCHRONOUSLY (a plain load + shared store would put one global-memory latency on
    // every step of the chain); fill-in rows (b < kl) and rows outside the matrix start as zeros.  One commit group per call.
    auto fetch = [&](i64 c) {
        if (c < n) {
            T *dst = colp(c);
            for (int b = tid; b < RP; b += NT) {
                const i64 i = c - kv + b;  // matrix row of band row b
                const bool ok = b >= kl && i >= 0 && i < m;
                const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + b);
                const int sz = ok ? (int)sizeof(T) : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(sa), "l"(ab + (ok ? b + c * ldab : 0)), "n"(sizeof(T)), "r"(sz) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (i64 c = 0; c < kv + 1 + TW_PF; ++c) fetch(c);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    i64 ju = 0;
    int info = 0;
    for (i64 j = 0; j < mn; ++j) {
        const int km = (int)(kl < m - 1 - j ? kl : m - 1 - j);
        T *cj = colp(j) + kv;  // cj[t] = A[j+t, j]
        if (tid < 32) {
            R best = (R)-1;
            int bidx = 0;
            for (int t = lane; t <= km; t += 32) {
                const R v = N::abs1(cj[t]);
                if (v > best) { best = v; bidx = t; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const R ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
            }
            if (lane == 0) {
                if (!(best >= (R)0) || bidx > km) bidx = 0;  // all-NaN column: IxAMAX returns the first index
                s_jp = bidx;
                s_piv = cj[bidx];
                ipiv[j] = j + bidx + 1;
            }
        }
        __syncthreads();
        const int jp = s_jp;
        const T piv = s_piv;
        const bool nz = !N::iszero(piv);
        if (nz) {
            const i64 jun = j + ku + jp < n - 1 ? j + ku + jp : n - 1;
            if (jun > ju) ju = jun;
        } else if (info == 0) info = (int)(j + 1);
        const int ncol = nz ? (int)(ju - j) + 1 : 0, total = (km + 1) * ncol;
        const T rinv = nz ? N::recip(piv) : N::zero();
        T nv[E];
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int e = tid + q * NT;
            if (e < total) {
                const int cc = e / (km + 1), t = e - cc * (km + 1);   // column j+cc, row j+t
                const T *pc = colp(j + cc) + (kv - cc);               // pc[t] = A[j+t, j+cc]
                const int ts = t == 0 ? jp : (t == jp ? 0 : t);       // row interchange j <-> j+jp
                const T v = pc[ts];
                if (cc == 0) nv[q] = t == 0 ? v : N::mul(v, rinv);
                else if (t == 0) nv[q] = v;
                else {
                    const T lt = N::mul(cj[ts], rinv), uc = pc[jp];
                    nv[q] = N::fma(N::neg(lt), uc, v);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int e = tid + q * NT;
            if (e < total) {
                const int cc = e / (km + 1), t = e - cc * (km + 1);
                colp(j + cc)[kv - cc + t] = nv[q];
            }
        }
        __syncthreads();
        // column j is final: retire it (in-matrix rows only), then its slot takes column j + kv + 1 + PF
        {
            const T *src = colp(j);
            for (int b = tid; b < RP; b += NT) {
                const i64 i = j - kv + b;
                if (i >= 0 && i < m) ab[b + j * ldab] = src[b];
            }
        }
        __syncthreads();
        fetch(j + kv + 1 + TW_PF);
        // the column fetched TW_PF steps ago is complete before anybody can touch it (first use: TW_PF + 1 steps after its fetch);
        // the barriers of the next step publish it
        asm volatile("cp.async.wait_group %0;" ::"n"(TW_PF - 1) : "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // columns mn .. n-1 of a wide matrix and whatever is still in the ring
    for (i64 c = mn; c < n && c <= mn + kv + TW_PF; ++c) {
        const T *src = colp(c);
        for (int b = tid; b < RP; b += NT) {
            const i64 i = c - kv + b;
            if (i >= 0 && i < m) ab[b + c * ldab] = src[b];
        }
    }
    if (tid == 0) d_info[0] = info;
}

// ---- gbtrs: one CTA per right-hand side; op = 0 'N', 1 'T', 2 'C' -----------------------------------------------------------
// The right-hand side lives in global memory (L2); every sweep step is: the owner of x[j] finishes it, a barrier, everyone
// subtracts its multiple from the entries in reach.  (A correctness-first kernel: the tuned pipelines exist for Float64 only.)
template <typename T>
__global__ void __launch_bounds__(256)
tgbtrs(int op, i64 n, i64 kl, i64 ku, const T *__restrict__ ab, i64 ldab, const i64 *__restrict__ ipiv, T *__restrict__ b0, i64 ldb)
{
    typedef Num<T> N;
    T *b = b0 + (i64)blockIdx.x * ldb;
    const i64 kv = kl + ku;
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ T s_x;
    if (op == 0) {
        // L: for j ascending: interchange, b[j+1..j+lm] -= b[j] * l
        if (kl > 0)
            for (i64 j = 0; j < n - 1; ++j) {
                const i64 lm = kl < n - 1 - j ? kl : n - 1 - j;
                if (tid == 0) {
                    const i64 p = ipiv[j] - 1;
                    const T t0 = b[p];
                    if (p != j) { b[p] = b[j]; b[j] = t0; }
                    s_x = t0;
                }
                __syncthreads();
                const T xj = s_x;
                const T *l = ab + kv + 1 + j * ldab;
                for (i64 t = tid; t < lm; t += nt) b[j + 1 + t] = N::fma(N::neg(xj), l[t], b[j + 1 + t]);
                __syncthreads();
            }
        // U: for j descending: b[j] /= U[j,j]; b[j-kv..j-1] -= b[j] * U[.,j]
        for (i64 j = n - 1; j >= 0; --j) {
            const T *col = ab + kv + j * ldab;  // U[j-d, j] = col[-d]
            if (tid == 0) { const T xj = N::mul(b[j], N::recip(col[0])); b[j] = xj; s_x = xj; }
            __syncthreads();
            const T xj = s_x;
            const i64 reach = kv < j ? kv : j;
            for (i64 d = 1 + tid; d <= reach; d += nt) b[j - d] = N::fma(N::neg(xj), col[-d], b[j - d]);
            __syncthreads();
        }
    } else {
        const bool cj = op == 2;
        // U^T (or U^H): for j ascending: b[j] = (b[j] - sum_d op(U[j-d,j]) b[j-d]) / op(U[j,j])
        for (i64 j = 0; j < n; ++j) {
            const T *col = ab + kv + j * ldab;
            const i64 reach = kv < j ? kv : j;
            T acc = N::zero();
            for (i64 d = 1 + tid; d <= reach; d += nt) acc = N::fma(cj ? N::conj(col[-d]) : col[-d], b[j - d], acc);
            // block reduction
            __shared__ T s_red[8];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = N::add(acc, N::shfl_xor(acc, o));
            if ((tid & 31) == 0) s_red[tid >> 5] = acc;
            __syncthreads();
            if (tid == 0) {
                T s = s_red[0];
                for (int w = 1; w < (nt >> 5); ++w) s = N::add(s, s_red[w]);
                const T dg = cj ? N::conj(col[0]) : col[0];
                b[j] = N::mul(N::add(b[j], N::neg(s)), N::recip(dg));
            }
            __syncthreads();
        }
        // L^T (or L^H): for j descending: b[j] -= sum_t op(l[t]) b[j+1+t]; then the interchange
        if (kl > 0)
            for (i64 j = n - 2; j >= 0; --j) {
                const i64 lm = kl < n - 1 - j ? kl : n - 1 - j;
                const T *l = ab + kv + 1 + j * ldab;
                T acc = N::zero();
                for (i64 t = tid; t < lm; t += nt) acc = N::fma(cj ? N::conj(l[t]) : l[t], b[j + 1 + t], acc);
                __shared__ T s_red2[8];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc = N::add(acc, N::shfl_xor(acc, o));
                if ((tid & 31) == 0) s_red2[tid >> 5] = acc;
                __syncthreads();
                if (tid == 0) {
                    T s = s_red2[0];
                    for (int w = 1; w < (nt >> 5); ++w) s = N::add(s, s_red2[w]);
                    const T v = N::add(b[j], N::neg(s));
                    const i64 p = ipiv[j] - 1;
                    if (p != j) { b[j] = b[p]; b[p] = v; } else b[j] = v;
                }
                __syncthreads();
            }
    }
}

// ---- tbmv: x <- op(T) x, T triangular band ('U': T[i,j] at a[(k+i-j) + j*lda], 'L': a[(i-j) + j*lda]); op = 0 'N', 1 'T', 2 'C' ----
// one thread per row of op(T); the result goes to `y` (scratch) and is copied back by the caller
template <typename T>
__global__ void __launch_bounds__(256)
ttbmv(int up, int op, int unit, i64 n, i64 k, const T *__restrict__ a, i64 lda, const T *__restrict__ x, T *__restrict__ y)
{
    typedef Num<T> N;
    const bool cj = op == 2;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const T dg = a[(up ? k : 0) + i * lda];
        T acc = unit ? x[i] : N::mul(cj ? N::conj(dg) : dg, x[i]);
        if (op == 0) {
            if (up) { const i64 j1 = i + k < n - 1 ? i + k : n - 1; for (i64 j = i + 1; j <= j1; ++j) acc = N::fma(a[(k + i - j) + j * lda], x[j], acc); }
            else { const i64 j0 = i - k > 0 ? i - k : 0; for (i64 j = i - 1; j >= j0; --j) acc = N::fma(a[(i - j) + j * lda], x[j], acc); }
        } else {  // row i of op(T) = column i of T, contiguous
            const T *col = a + i * lda + (up ? k : 0);
            if (up) { const i64 dmax = i < k ? i : k; for (i64 d = 1; d <= dmax; ++d) acc = N::fma(cj ? N::conj(col[-d]) : col[-d], x[i - d], acc); }
            else { const i64 dmax = n - 1 - i < k ? n - 1 - i : k; for (i64 d = 1; d <= dmax; ++d) acc = N::fma(cj ? N::conj(col[d]) : col[d], x[i + d], acc); }
        }
        y[i] = acc;
    }
}

// ---- tbsv: x <- op(T)^{-1} x; one CTA per right-hand side (column of b0, stride ldb); correctness-first chain ----
template <typename T>
__global__ void __launch_bounds__(256)
ttbsv(int up, int op, int unit, i64 n, i64 k, const T *__restrict__ a, i64 lda, T *__restrict__ b0, i64 ldb)
{
    typedef Num<T> N;
    T *x = b0 + (i64)blockIdx.x * ldb;
    const int tid = threadIdx.x, nt = blockDim.x;
    const bool cj = op == 2;
    __shared__ T s_x;
    __shared__ T s_red[8];
    const bool ascending = (op == 0) ? !up : up;  // order in which the unknowns become known
    for (i64 c = 0; c < n; ++c) {
        const i64 j = ascending ? c : n - 1 - c;
        const T *col = a + j * lda + (up ? k : 0);  // diagonal entry of column j; T[j-d, j] = col[-d] ('U'), T[j+d, j] = col[+d] ('L')
        if (op == 0) {
            // column sweep: x[j] /= T[j,j]; the entries in reach get -x[j] * T[., j]
            if (tid == 0) { const T xj = unit ? x[j] : N::mul(x[j], N::recip(col[0])); x[j] = xj; s_x = xj; }
            __syncthreads();
            const T xj = s_x;
            const i64 reach = up ? (k < j ? k : j) : (k < n - 1 - j ? k : n - 1 - j);
            for (i64 d = 1 + tid; d <= reach; d += nt) {
                const i64 i = up ? j - d : j + d;
                x[i] = N::fma(N::neg(xj), up ? col[-d] : col[d], x[i]);
            }
            __syncthreads();
        } else {
            // dot-product sweep: x[j] = (x[j] - sum_d op(T[j-+d, j]) x[j-+d]) / op(T[j,j])
            const i64 reach = up ? (k < j ? k : j) : (k < n - 1 - j ? k : n - 1 - j);
            T acc = N::zero();
            for (i64 d = 1 + tid; d <= reach; d += nt) {
                const T t = up ? col[-d] : col[d];
                acc = N::fma(cj ? N::conj(t) : t, x[up ? j - d : j + d], acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = N::add(acc, N::shfl_xor(acc, o));
            if ((tid & 31) == 0) s_red[tid >> 5] = acc;
            __syncthreads();
            if (tid == 0) {
                T sum = s_red[0];
                for (int w = 1; w < (nt >> 5); ++w) sum = N::add(sum, s_red[w]);
                T v = N::add(x[j], N::neg(sum));
                if (!unit) v = N::mul(v, N::recip(cj ? N::conj(col[0]) : col[0]));
                x[j] = v;
            }
            __syncthreads();
        }
    }
}

// ---- pbtf2: unblocked Hermitian band Cholesky in xPBTF2's order; one CTA; up: A = U^H U, else A = L L^H ----
template <typename T>
__global__ void __launch_bounds__(256)
tpbtf2(int up, i64 n, i64 kd, T *__restrict__ ab, i64 ldab, int *__restrict__ d_info)
{
    typedef Num<T> N;
    typedef typename N::real R;
    const int tid = threadIdx.x, nt = blockDim.x;
    // S(i, c) for i <= c: 'U' ab[(kd + i - c) + c*ldab]; 'L' holds its conjugate at ab[(c - i) + i*ldab]
    for (i64 j = 0; j < n; ++j) {
        T *dj = ab + (up ? kd : 0) + j * ldab;
        const R ajj = reinterpret_cast<const R *>(dj)[0];  // real part of the diagonal entry (xPBTF2 reads nothing else of it)
        if (!(ajj > (R)0)) {  // xPBTF2: AJJ <= 0 (NaN continues in LAPACK; a NaN pivot makes everything NaN either way)
            if (ajj <= (R)0) { if (tid == 0) d_info[0] = (int)(j + 1); return; }
        }
        const R d = sqrt(ajj), rinv = (R)1 / d;
        const i64 kn = kd < n - 1 - j ? kd : n - 1 - j;
        __syncthreads();  // everyone has read the pivot
        if (tid == 0) { reinterpret_cast<R *>(dj)[0] = d; if (sizeof(T) != sizeof(R)) reinterpret_cast<R *>(dj)[1] = (R)0; }
        // scale row j of U ('U') / column j of L ('L') by 1/d
        for (i64 c = 1 + tid; c <= kn; c += nt) {
            T *e = up ? ab + (kd - c) + (j + c) * ldab : ab + c + j * ldab;
            R *er = reinterpret_cast<R *>(e);
            er[0] *= rinv;
            if (sizeof(T) != sizeof(R)) er[1] *= rinv;
        }
        __syncthreads();
        // trailing update: 'U': S(j+r, j+c) -= conj(u_r) u_c (r <= c), u = row j;  'L': L(j+c, j+r) -= l_c conj(l_r) (c >= r), l = column j
        const i64 total = kn * (kn + 1) / 2;
        for (i64 e = tid; e < total; e += nt) {
            i64 c = 1;
            while (c * (c + 1) / 2 <= e) ++c;
            const i64 r = e - c * (c - 1) / 2 + 1;
            if (up) {
                const T ur = ab[(kd - r) + (j + r) * ldab], uc = ab[(kd - c) + (j + c) * ldab];
                T *t = ab + (kd + r - c) + (j + c) * ldab;
                *t = N::fma(N::neg(N::conj(ur)), uc, *t);
                if (r == c) *t = N::realpart(*t);
            } else {
                const T lr = ab[r + j * ldab], lc = ab[c + j * ldab];
                T *t = ab + (c - r) + (j + r) * ldab;
                *t = N::fma(N::neg(lc), N::conj(lr), *t);
                if (r == c) *t = N::realpart(*t);
            }
        }
        __syncthreads();
    }
}

// ---- banded x banded, banded x dense for the other element types: one thread per entry of C, inner index ascending ----
template <typename T>
__global__ void __launch_bounds__(256)
tgbmm_bb(i64 n, i64 nu, i64 m, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu, T alpha, const T *__restrict__ a, i64 lda, const T *__restrict__ b,
         i64 ldb, T beta, T *__restrict__ c, i64 ldc)
{
    typedef Num<T> N;
    const i64 W = Cl + Cu + 1, total = W * m;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 j = e / W, r = e - j * W, k = j + r - Cu;
        if (k < 0 || k >= n) continue;
        T *cp = c + r + j * ldc;
        T acc = N::iszero(beta) ? N::zero() : N::mul(beta, *cp);
        if (!N::iszero(alpha)) {
            i64 v0 = k - Al > j - Bu ? k - Al : j - Bu, v1 = k + Au < j + Bl ? k + Au : j + Bl;
            if (v0 < 0) v0 = 0;
            if (v1 > nu - 1) v1 = nu - 1;
            for (i64 v = v0; v <= v1; ++v) acc = N::fma(N::mul(alpha, b[(Bu + v - j) + j * ldb]), a[(Au + k - v) + v * lda], acc);
        }
        *cp = acc;
    }
}
template <typename T>
__global__ void __launch_bounds__(256)
tgbmm_bd(int op, i64 m, i64 n, i64 kl, i64 ku, i64 nrhs, T alpha, const T *__restrict__ a, i64 lda, const T *__restrict__ b, i64 ldb, T beta,
         T *__restrict__ c, i64 ldc)
{
    typedef Num<T> N;
    const i64 rows = op == 0 ? m : n, total = rows * nrhs;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 q = e / rows, i = e - q * rows;
        T *cp = c + i + q * ldc;
        const T *bq = b + q * ldb;
        T acc = N::iszero(beta) ? N::zero() : N::mul(beta, *cp);
        if (!N::iszero(alpha)) {
            if (op == 0) {
                const i64 j0 = i - kl > 0 ? i - kl : 0, j1 = i + ku < n - 1 ? i + ku : n - 1;
                for (i64 j = j0; j <= j1; ++j) acc = N::fma(N::mul(alpha, bq[j]), a[(ku + i - j) + j * lda], acc);
            } else {  // row i of op(A) = column i of A
                const i64 r0 = i - ku > 0 ? i - ku : 0, r1 = i + kl < m - 1 ? i + kl : m - 1;
                const T *col = a + i * lda + (ku - i);
                T t = N::zero();
                for (i64 r = r0; r <= r1; ++r) t = N::fma(op == 2 ? N::conj(col[r]) : col[r], bq[r], t);
                acc = N::fma(alpha, t, acc);
            }
        }
        *cp = acc;
    }
}

template <typename T> __host__ T host_scalar(const void *p);
template <> __host__ float host_scalar<float>(const void *p) { return *(const float *)p; }
template <> __host__ double host_scalar<double>(const void *p) { return *(const double *)p; }
template <> __host__ cuFloatComplex host_scalar<cuFloatComplex>(const void *p) { const float *q = (const float *)p; return make_cuFloatComplex(q[0], q[1]); }
template <> __host__ cuDoubleComplex host_scalar<cuDoubleComplex>(const void *p) { const double *q = (const double *)p; return make_cuDoubleComplex(q[0], q[1]); }

template <typename T>
int gbmv_impl(bmb200_ctx *h, char trans, i64 m, i64 n, i64 kl, i64 ku, const void *alpha, const void *dA, i64 lda, const void *dx, i64 incx,
              const void *beta, void *dy, i64 incy)
{
    if (!h) return -1;
    const bool tn = trans == 'N' || trans == 'n', tt = trans == 'T' || trans == 't', tc = trans == 'C' || trans == 'c';
    if (!tn && !tt && !tc) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (kl < 0) return -5;
    if (ku < 0) return -6;
    if (!alpha) return -7;
    if (lda < kl + ku + 1) return -9;
    if (incx == 0) return -11;
    if (!beta) return -12;
    if (incy == 0) return -14;
    const i64 leny = tn ? m : n;
    if (m == 0 || n == 0 || leny == 0) return 0;
    if (!dA || !dx || !dy) return -8;
    DeviceGuard g(h->device);
    const T al = host_scalar<T>(alpha), be = host_scalar<T>(beta);
    const T *A = (const T *)dA, *x = (const T *)dx;
    T *y = (T *)dy;
    if (tn) {
        const i64 blocks = imin64(cdiv64(m, 256), (i64)h->sm_count * 16);
        tgbmv_n<T><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, al, A, lda, x, incx, be, y, incy);
    } else if (kl + ku + 1 <= 24) {
        const i64 blocks = imin64(cdiv64(n, 256), (i64)h->sm_count * 16);
        if (tc) tgbmv_t_thread<T, true><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, al, A, lda, x, incx, be, y, incy);
        else tgbmv_t_thread<T, false><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, al, A, lda, x, incx, be, y, incy);
    } else {
        const i64 blocks = imin64(cdiv64(n, 8), (i64)h->sm_count * 16);
        if (tc) tgbmv_t<T, true><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, al, A, lda, x, incx, be, y, incy);
        else tgbmv_t<T, false><<<(unsigned)blocks, 256, 0, h->stream>>>(m, n, kl, ku, al, A, lda, x, incx, be, y, incy);
    }
    BMB_LAUNCH_CHECK(h);
    return 0;
}

template <typename T>
int hbmv_impl(bmb200_ctx *h, char uplo, i64 n, i64 k, const void *alpha, const void *dA, i64 lda, const void *dx, i64 incx, const void *beta,
              void *dy, i64 incy)
{
    if (!h) return -1;
    const int up = (uplo == 'U' || uplo == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -2;
    if (n < 0) return -3;
    if (k < 0) return -4;
    if (!alpha) return -5;
    if (lda < k + 1) return -7;
    if (incx != 1) return -9;
    if (!beta) return -10;
    if (incy != 1) return -12;
    if (n == 0) return 0;
    if (!dA || !dx || !dy) return -6;
    if (dx == dy) return -8;
    DeviceGuard g(h->device);
    const i64 blocks = imin64(cdiv64(n, 256), (i64)h->sm_count * 16);
    thbmv<T><<<(unsigned)blocks, 256, 0, h->stream>>>(up, n, k, host_scalar<T>(alpha), (const T *)dA, lda, (const T *)dx, host_scalar<T>(beta), (T *)dy);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

template <typename T>
int gbtrf_impl(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, void *dAB, i64 ldab, i64 *d_ipiv, int *info)
{
    if (!h) return -1;
    if (m < 0) return -2;
    if (n < 0) return -3;
    if (kl < 0) return -4;
    if (ku < 0) return -5;
    if (ldab < 2 * kl + ku + 1) return -7;
    if (!info) return -9;
    *info = 0;
    if (m == 0 || n == 0) return 0;
    if (!dAB) return -6;
    if (!d_ipiv) return -8;
    DeviceGuard g(h->device);
    int *d_info = h->d_info + 28;
    const i64 work = (kl + 1) * (kl + ku + 1);
    // narrow bands: the shared-memory window kernel (the live columns never leave the SM); otherwise the global-memory kernel
    int slots = 8;
    while (slots < kl + ku + 1 + TW_PF + 1) slots <<= 1;
    const int RP = (int)(2 * kl + ku + 1);
    const size_t smem = (size_t)slots * RP * sizeof(T);
    // (measured, n = 10^5 / 2*10^4, ns per column, window / global: (4,3) 1078 / 1308, (32,32) 2802 / 2277: the window kernel pays
    // for its per-entry index arithmetic once a thread owns several entries, so it only takes the narrowest bands)
    if (work <= 256 && smem <= 200 * 1024 && !h->tune.typed_nowin) {
        auto launch = [&](auto kern, unsigned nt) -> int {
            BMB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<1, nt, smem, h->stream>>>(m, n, (int)kl, (int)ku, (T *)dAB, ldab, d_ipiv, d_info, slots, RP);
            return 0;
        };
        int rc;
        if (work <= 64) rc = launch(tgbtf2_win<T, 64, 1>, 64);
        else rc = launch(tgbtf2_win<T, 128, 2>, 128);
        if (rc) return rc;
    } else {
        const unsigned nt = work <= 64 ? 64 : (work <= 1024 ? 256 : 1024);
        tgbtf2<T><<<1, nt, 0, h->stream>>>(m, n, kl, ku, (T *)dAB, ldab, d_ipiv, d_info);
    }
    BMB_LAUNCH_CHECK(h);
    BMB_CUDA(h, cudaMemcpyAsync(info, d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int gbtrs_impl(bmb200_ctx *h, char trans, i64 n, i64 kl, i64 ku, i64 nrhs, const void *dAB, i64 ldab, const i64 *d_ipiv, void *dB, i64 ldb)
{
    if (!h) return -1;
    const int op = (trans == 'N' || trans == 'n') ? 0 : ((trans == 'T' || trans == 't') ? 1 : ((trans == 'C' || trans == 'c') ? 2 : -1));
    if (op < 0) return -2;
    if (n < 0) return -3;
    if (kl < 0) return -4;
    if (ku < 0) return -5;
    if (nrhs < 0) return -6;
    if (ldab < 2 * kl + ku + 1) return -8;
    if (ldb < (n > 1 ? n : 1)) return -11;
    if (n == 0 || nrhs == 0) return 0;
    if (!dAB || !d_ipiv || !dB) return -7;
    DeviceGuard g(h->device);
    tgbtrs<T><<<(unsigned)nrhs, 256, 0, h->stream>>>(op, n, kl, ku, (const T *)dAB, ldab, d_ipiv, (T *)dB, ldb);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

static int tb_op(char trans) { return (trans == 'N' || trans == 'n') ? 0 : ((trans == 'T' || trans == 't') ? 1 : ((trans == 'C' || trans == 'c') ? 2 : -1)); }

template <typename T>
int tb_impl(bmb200_ctx *h, bool solve, char uplo, char trans, char diag, i64 n, i64 k, const void *dA, i64 lda, void *dx, i64 incx)
{
    if (!h) return -1;
    const int up = (uplo == 'U' || uplo == 'u'), unit = (diag == 'U' || diag == 'u'), op = tb_op(trans);
    if (!up && !(uplo == 'L' || uplo == 'l')) return -2;
    if (op < 0) return -3;
    if (!unit && !(diag == 'N' || diag == 'n')) return -4;
    if (n < 0) return -5;
    if (k < 0) return -6;
    if (lda < k + 1) return -8;
    if (incx != 1) return -10;
    if (n == 0) return 0;
    if (!dA || !dx) return -7;
    DeviceGuard g(h->device);
    if (solve) {
        ttbsv<T><<<1, 256, 0, h->stream>>>(up, op, unit, n, k, (const T *)dA, lda, (T *)dx, n);
        BMB_LAUNCH_CHECK(h);
        return 0;
    }
    if (bmb_ensure_scratch(h, (size_t)n * sizeof(T)) != 0) return BMB200_ERR_CUDA;
    const i64 blocks = imin64(cdiv64(n, 256), (i64)h->sm_count * 16);
    ttbmv<T><<<(unsigned)blocks, 256, 0, h->stream>>>(up, op, unit, n, k, (const T *)dA, lda, (const T *)dx, (T *)h->scratch);
    BMB_LAUNCH_CHECK(h);
    BMB_CUDA(h, cudaMemcpyAsync(dx, h->scratch, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

template <typename T>
int pbtrf_impl(bmb200_ctx *h, char uplo, i64 n, i64 kd, void *dAB, i64 ldab, int *info)
{
    if (!h) return -1;
    const int up = (uplo == 'U' || uplo == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -2;
    if (n < 0) return -3;
    if (kd < 0) return -4;
    if (ldab < kd + 1) return -6;
    if (!info) return -7;
    *info = 0;
    if (n == 0) return 0;
    if (!dAB) return -5;
    DeviceGuard g(h->device);
    int *d_info = h->d_info + 28;
    BMB_CUDA(h, cudaMemsetAsync(d_info, 0, sizeof(int), h->stream));
    tpbtf2<T><<<1, 256, 0, h->stream>>>(up, n, kd > n - 1 ? (n > 1 ? n - 1 : 0) : kd, (T *)dAB + (up ? kd - (kd > n - 1 ? (n > 1 ? n - 1 : 0) : kd) : 0), ldab, d_info);
    BMB_LAUNCH_CHECK(h);
    BMB_CUDA(h, cudaMemcpyAsync(info, d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename T>
int pbtrs_impl(bmb200_ctx *h, char uplo, i64 n, i64 kd, i64 nrhs, const void *dAB, i64 ldab, void *dB, i64 ldb)
{
    if (!h) return -1;
    const int up = (uplo == 'U' || uplo == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -2;
    if (n < 0) return -3;
    if (kd < 0) return -4;
    if (nrhs < 0) return -5;
    if (ldab < kd + 1) return -7;
    if (ldb < (n > 1 ? n : 1)) return -9;
    if (n == 0 || nrhs == 0) return 0;
    if (!dAB || !dB) return -6;
    DeviceGuard g(h->device);
    // xPBTRS: 'U': U^H y = b, U x = y;  'L': L y = b, L^H x = y
    ttbsv<T><<<(unsigned)nrhs, 256, 0, h->stream>>>(up, up ? 2 : 0, 0, n, kd, (const T *)dAB, ldab, (T *)dB, ldb);
    BMB_LAUNCH_CHECK(h);
    ttbsv<T><<<(unsigned)nrhs, 256, 0, h->stream>>>(up, up ? 0 : 2, 0, n, kd, (const T *)dAB, ldab, (T *)dB, ldb);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

template <typename T>
int gbmm_bb_impl(bmb200_ctx *h, i64 n, i64 nu, i64 m, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu, const void *alpha, const void *dA, i64 lda,
                 const void *dB, i64 ldb, const void *beta, void *dC, i64 ldc)
{
    if (!h) return -1;
    if (n < 0) return -2;
    if (nu < 0) return -3;
    if (m < 0) return -4;
    if (Al < 0) return -5;
    if (Au < 0) return -6;
    if (Bl < 0) return -7;
    if (Bu < 0) return -8;
    if (Cl < 0 || Cl > Al + Bl) return -9;
    if (Cu < 0 || Cu > Au + Bu) return -10;
    if (!alpha) return -11;
    if (lda < Al + Au + 1) return -13;
    if (ldb < Bl + Bu + 1) return -15;
    if (!beta) return -16;
    if (ldc < Cl + Cu + 1) return -18;
    if (n == 0 || m == 0) return 0;
    if (!dC || (nu > 0 && (!dA || !dB))) return -12;
    DeviceGuard g(h->device);
    const i64 blocks = imin64(cdiv64((Cl + Cu + 1) * m, 256), (i64)h->sm_count * 16);
    tgbmm_bb<T><<<(unsigned)blocks, 256, 0, h->stream>>>(n, nu, m, Al, Au, Bl, Bu, Cl, Cu, host_scalar<T>(alpha), (const T *)dA, lda, (const T *)dB, ldb,
                                                         host_scalar<T>(beta), (T *)dC, ldc);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

template <typename T>
int gbmm_bd_impl(bmb200_ctx *h, char trans, i64 m, i64 n, i64 kl, i64 ku, i64 nrhs, const void *alpha, const void *dA, i64 lda, const void *dB, i64 ldb,
                 const void *beta, void *dC, i64 ldc)
{
    if (!h) return -1;
    const int op = tb_op(trans);
    if (op < 0) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (kl < 0) return -5;
    if (ku < 0) return -6;
    if (nrhs < 0) return -7;
    if (!alpha) return -8;
    if (lda < kl + ku + 1) return -10;
    const i64 rowsB = op ? m : n, rowsC = op ? n : m;
    if (ldb < imax64(1, rowsB)) return -12;
    if (!beta) return -13;
    if (ldc < imax64(1, rowsC)) return -15;
    if (rowsC == 0 || nrhs == 0) return 0;
    if (!dC || (rowsB > 0 && (!dA || !dB))) return -9;
    DeviceGuard g(h->device);
    const i64 blocks = imin64(cdiv64(rowsC * nrhs, 256), (i64)h->sm_count * 16);
    tgbmm_bd<T><<<(unsigned)blocks, 256, 0, h->stream>>>(op, m, n, kl, ku, nrhs, host_scalar<T>(alpha), (const T *)dA, lda, (const T *)dB, ldb,
                                                         host_scalar<T>(beta), (T *)dC, ldc);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

}  // namespace

#define BMB_TYPED_EXPORTS(P, T)                                                                                                                   \
    extern "C" int bmb200_##P##gbmv(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, const void *alpha, const void *dA, \
                                    int64_t lda, const void *dx, int64_t incx, const void *beta, void *dy, int64_t incy)                          \
    {                                                                                                                                             \
        return gbmv_impl<T>(h, trans, m, n, kl, ku, alpha, dA, lda, dx, incx, beta, dy, incy);                                                    \
    }                                                                                                                                             \
    extern "C" int bmb200_##P##gbtrf(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, void *dAB, int64_t ldab, int64_t *d_ipiv,     \
                                     int *info)                                                                                                   \
    {                                                                                                                                             \
        return gbtrf_impl<T>(h, m, n, kl, ku, dAB, ldab, d_ipiv, info);                                                                           \
    }                                                                                                                                             \
    extern "C" int bmb200_##P##gbtrs(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs, const void *dAB, int64_t ldab, \
                                     const int64_t *d_ipiv, void *dB, int64_t ldb)                                                                \
    {                                                                                                                                             \
        return gbtrs_impl<T>(h, trans, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);                                                              \
    }

BMB_TYPED_EXPORTS(s, float)
BMB_TYPED_EXPORTS(c, cuFloatComplex)
BMB_TYPED_EXPORTS(z, cuDoubleComplex)

#define BMB_TYPED_EXPORTS2(P, T)                                                                                                                     \
    extern "C" int bmb200_##P##tbsv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda, void *dx,  \
                                    int64_t incx)                                                                                                    \
    {                                                                                                                                                \
        return tb_impl<T>(h, true, uplo, trans, diag, n, k, dA, lda, dx, incx);                                                                      \
    }                                                                                                                                                \
    extern "C" int bmb200_##P##tbmv(bmb200_handle_t h, char uplo, char trans, char diag, int64_t n, int64_t k, const void *dA, int64_t lda, void *dx,  \
                                    int64_t incx)                                                                                                    \
    {                                                                                                                                                \
        return tb_impl<T>(h, false, uplo, trans, diag, n, k, dA, lda, dx, incx);                                                                     \
    }                                                                                                                                                \
    extern "C" int bmb200_##P##pbtrf(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, void *dAB, int64_t ldab, int *info)                          \
    {                                                                                                                                                \
        return pbtrf_impl<T>(h, uplo, n, kd, dAB, ldab, info);                                                                                       \
    }                                                                                                                                                \
    extern "C" int bmb200_##P##pbtrs(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, int64_t nrhs, const void *dAB, int64_t ldab, void *dB,       \
                                     int64_t ldb)                                                                                                    \
    {                                                                                                                                                \
        return pbtrs_impl<T>(h, uplo, n, kd, nrhs, dAB, ldab, dB, ldb);                                                                              \
    }                                                                                                                                                \
    extern "C" int bmb200_##P##gbmm_bb(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au, int64_t Bl, int64_t Bu, int64_t Cl, \
                                       int64_t Cu, const void *alpha, const void *dA, int64_t lda, const void *dB, int64_t ldb, const void *beta,     \
                                       void *dC, int64_t ldc)                                                                                        \
    {                                                                                                                                                \
        return gbmm_bb_impl<T>(h, n, nu, m, Al, Au, Bl, Bu, Cl, Cu, alpha, dA, lda, dB, ldb, beta, dC, ldc);                                         \
    }                                                                                                                                                \
    extern "C" int bmb200_##P##gbmm_bd(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku, int64_t nrhs, const void *alpha,   \
                                       const void *dA, int64_t lda, const void *dB, int64_t ldb, const void *beta, void *dC, int64_t ldc)             \
    {                                                                                                                                                \
        return gbmm_bd_impl<T>(h, trans, m, n, kl, ku, nrhs, alpha, dA, lda, dB, ldb, beta, dC, ldc);                                                \
    }
BMB_TYPED_EXPORTS2(s, float)
BMB_TYPED_EXPORTS2(c, cuFloatComplex)
BMB_TYPED_EXPORTS2(z, cuDoubleComplex)

// the generic kernels instantiated for double: a cross-check of the tuned Float64 path (tests), not a dispatch target
extern "C" int bmb200_internal_dgbtrf_generic(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double *dAB, int64_t ldab, int64_t *d_ipiv, int *info)
{
    return gbtrf_impl<double>(h, m, n, kl, ku, dAB, ldab, d_ipiv, info);
}
extern "C" int bmb200_internal_dgbtrs_generic(bmb200_handle_t h, char trans, int64_t n, int64_t kl, int64_t ku, int64_t nrhs, const double *dAB, int64_t ldab,
                                              const int64_t *d_ipiv, double *dB, int64_t ldb)
{
    return gbtrs_impl<double>(h, trans, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
}

extern "C" int bmb200_ssbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, const void *alpha, const void *dA, int64_t lda, const void *dx, int64_t incx,
                            const void *beta, void *dy, int64_t incy)
{
    return hbmv_impl<float>(h, uplo, n, k, alpha, dA, lda, dx, incx, beta, dy, incy);
}
extern "C" int bmb200_chbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, const void *alpha, const void *dA, int64_t lda, const void *dx, int64_t incx,
                            const void *beta, void *dy, int64_t incy)
{
    return hbmv_impl<cuFloatComplex>(h, uplo, n, k, alpha, dA, lda, dx, incx, beta, dy, incy);
}
extern "C" int bmb200_zhbmv(bmb200_handle_t h, char uplo, int64_t n, int64_t k, const void *alpha, const void *dA, int64_t lda, const void *dx, int64_t incx,
                            const void *beta, void *dy, int64_t incy)
{
    return hbmv_impl<cuDoubleComplex>(h, uplo, n, k, alpha, dA, lda, dx, incx, beta, dy, incy);
}
