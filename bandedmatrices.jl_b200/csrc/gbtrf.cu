// gbtrf.cu -- partial-pivot LU of a general band matrix in LAPACK LU storage, on sm_100a.
//
// Replaces dgbtrf_ (LAPACK.gbtrf! at src/banded/BandedLU.jl:98).  Contract kept from the reference
// CPU path: ipiv = FIRST maximum of |.| over rows j..j+kl of the current column (IDAMAX),
// multipliers = column * (1/pivot) (DSCAL with the reciprocal), trailing update one FMA per
// element per eliminated column in ascending column order (DGER), fill-in rows zeroed, multipliers
// left un-permuted.  Any schedule that preserves the per-element order of those FMAs gives the same
// bits as DGBTF2, so pivots AND factors are bit-identical to the reference for every band width.
//
// Kernels
//   gbtrf_window   one persistent CTA; the active (kl+1) x (kl+ku+1) window lives in shared memory
//                  as a ring of column slots fed by cp.async PF columns ahead; used while the ring
//                  fits in 227 KB.  The factorisation is a chain of min(m,n) dependent pivot steps
//                  (no parallel-in-n algorithm keeps ipiv), so the figure of
//                  merit is ns per column, not a roofline fraction.
//   (wide bands: gbtrf_blocked.cu)
#include "common.cuh"

int bmb_gbtrf_blocked(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv);
int bmb_gbtrf_reg(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv);  // gbtrf_reg.cu
int bmb_gbtrf_mw(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv);   // gbtrf_mw.cu

#define GBTRF_PF 12

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Fetch column c of AB into its ring slot: band rows from global, fill-in rows [max(0,kv-c), kl) zero.
__device__ __forceinline__ void fetch_column(double *slot, const double *__restrict__ ab, i64 ldab, i64 c, int ldw,
                                             int kl, int kv)
{
    const i64 z0 = (i64)kv - c;  // first zeroed row (may be <= 0)
    for (int r = threadIdx.x; r < ldw; r += blockDim.x) {
        if (r < kl && r >= z0) slot[r] = 0.0;
        else cp_async8(slot + r, ab + r + c * ldab);
    }
}

__global__ void __launch_bounds__(1024, 1)
gbtrf_window(i64 m, i64 n, int kl, int ku, double *__restrict__ ab, i64 ldab, i64 *__restrict__ ipiv,
             int *__restrict__ d_info, int nslot)
{
    extern __shared__ double sm[];
    const int kv = kl + ku, ldw = kl + kv + 1;
    double *urow = sm + (size_t)nslot * ldw;  // kv+1 : old pivot row
    double *row0 = urow + (kv + 1);           // kv+1 : old row j
    double *lcol = row0 + (kv + 1);           // kl+1 : scaled multipliers
    __shared__ int s_jp;
    __shared__ int s_info;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nthr = blockDim.x;
    // fixed (row, column-group) coordinates of this thread inside the window: no div/mod in the loop
    const int ti = tid % (kl + 1), tc = tid / (kl + 1), tcs = nthr / (kl + 1);
    const bool tact = tc < tcs;  // threads beyond the last full column group idle in (4b)
    const i64 mn = m < n ? m : n;
    const i64 ncols = (mn + kv < n) ? mn + kv : n;  // columns the factorisation can touch
#define SLOTI(s) (sm + (size_t)(s) * ldw)

    if (tid == 0) s_info = 0;
    // initial window: columns 0 .. kv+PF-1 (slot of column c is c mod nslot; nslot > kv+PF)
    for (i64 c = 0; c < ncols && c < kv + GBTRF_PF; ++c) fetch_column(SLOTI((int)c), ab, ldab, c, ldw, kl, kv);
    cp_async_commit();
    cp_async_wait<0>();

    i64 ju = 0;   // 0-based last column touched so far
    int sj = 0;   // slot of column j
    for (i64 j = 0; j < mn; ++j) {
        const int km = (int)((kl < m - 1 - j) ? kl : (m - 1 - j));
        {   // (1) prefetch column j+kv+PF (its slot held column j-2, already written back)
            const i64 cf = j + kv + GBTRF_PF;
            int sf = sj + kv + GBTRF_PF;
            if (sf >= nslot) sf -= nslot;
            if (cf < ncols) fetch_column(SLOTI(sf), ab, ldab, cf, ldw, kl, kv);
            cp_async_commit();
        }
        cp_async_wait<GBTRF_PF>();
        __syncthreads();  // (A) column j+kv landed; step j-1's updates are visible
        double *colj = SLOTI(sj);
        if (wid == 0) {  // (3) IDAMAX over rows j..j+km of column j: first maximum
            double best = -1.0;
            int bidx = 0;
            for (int i = lane; i <= km; i += 32) {
                const double v = fabs(colj[kv + i]);
                if (v > best) { best = v; bidx = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            if (lane == 0) {
                s_jp = bidx;
                ipiv[j] = j + bidx + 1;
            }
        } else if (j > 0) {  // write back the finished column j-1 while warp 0 searches
            const double *src = SLOTI(sj == 0 ? nslot - 1 : sj - 1);
            double *dst = ab + (j - 1) * ldab;
            for (int r = tid - 32; r < ldw; r += nthr - 32) dst[r] = src[r];
        }
        __syncthreads();  // (B)
        const int jp = s_jp;
        const double pv = colj[kv + jp];
        if (pv != 0.0) {
            i64 cand = j + ku + jp;
            if (cand > n - 1) cand = n - 1;
            if (cand > ju) ju = cand;
            const int nc = (int)(ju - j);  // columns right of j that are touched
            const double rinv = 1.0 / pv;
            // (4a) stage the old pivot row, the old row j and the scaled multiplier column
            for (int t = tid; t <= nc; t += nthr) {
                int st = sj + t;
                if (st >= nslot) st -= nslot;
                const double *s = SLOTI(st);
                urow[t] = s[kv + jp - t];
                row0[t] = s[kv - t];
            }
            for (int i = nthr - 1 - tid; i <= km; i += nthr) {  // from the far end: other warps than (4a)'s first loop
                const double v = (i == jp) ? colj[kv] : colj[kv + i];
                lcol[i] = (i == 0) ? pv : __dmul_rn(v, rinv);
            }
            __syncthreads();  // (C)
            // (4b) swap + scale + rank-1 update, every element touched by exactly one thread
            if (tact && ti <= km) {
                const double li = lcol[ti];
                for (int c = tc; c <= nc; c += tcs) {
                    int sc = sj + c;
                    if (sc >= nslot) sc -= nslot;
                    double *s = SLOTI(sc);
                    if (ti == 0) s[kv - c] = urow[c];          // pivot row moves up (column 0: the pivot itself)
                    else if (c == 0) s[kv + ti] = li;          // multipliers
                    else {
                        const double aold = (ti == jp) ? row0[c] : s[kv + ti - c];
                        s[kv + ti - c] = fma(-urow[c], li, aold);
                    }
                }
            }
        } else {
            if (tid == 0 && s_info == 0) s_info = (int)(j + 1);
        }
        // the barrier (A) of the next step orders (4b) before the next pivot search
        sj = (sj + 1 == nslot) ? 0 : sj + 1;
    }
    cp_async_wait<0>();
    __syncthreads();
    // write back the last finished column and the touched-but-unfinished ones (n > m only)
    for (i64 c = (mn > 0 ? mn - 1 : 0); c < ncols && c <= mn - 1 + kv; ++c) {
        const double *src = SLOTI((int)(c % nslot));
        double *dst = ab + c * ldab;
        for (int r = tid; r < ldw; r += nthr) dst[r] = src[r];
    }
    if (tid == 0) d_info[0] = s_info;
#undef SLOTI
}

// ------------------------------------------------------------------------------------------------

extern "C" int bmb200_dgbtrf(bmb200_handle_t h, int64_t m, int64_t n, int64_t kl, int64_t ku, double *dAB,
                             int64_t ldab, int64_t *d_ipiv, int *info)
{
    if (!h) return -1;
    if (m < 0) return -2;
    if (n < 0) return -3;
    if (kl < 0) return -4;
    if (ku < 0) return -5;
    if (ldab < 2 * kl + ku + 1) return -7;
    if (info) *info = 0;
    if (m == 0 || n == 0) return 0;
    if (!dAB) return -6;
    if (!d_ipiv) return -8;
    DeviceGuard g(h->device);
    // band widths beyond the matrix cannot hold data; clamp so the window stays small for tiny matrices
    const i64 kv = kl + ku, ldw = kl + kv + 1;
    const int nslot = (int)(kv + GBTRF_PF + 2);
    const size_t smem = ((size_t)nslot * ldw + 2 * (kv + 1) + (kl + 1)) * sizeof(double);
    int rc;
    if (kl <= 31 && kv + 1 <= 33) {  // register-window kernels: several warps (gbtrf_mw.cu), else one (gbtrf_reg.cu)
        rc = h->tune.gbtrf_nomw ? 1 : bmb_gbtrf_mw(h, m, n, kl, ku, dAB, ldab, d_ipiv);
        if (rc == 1) rc = bmb_gbtrf_reg(h, m, n, kl, ku, dAB, ldab, d_ipiv);
        if (rc) return rc;
    } else if (smem <= 220 * 1024) {
        BMB_CUDA(h, cudaFuncSetAttribute(gbtrf_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // (kl+1) x G threads, G column groups: one (row, column) element per thread when the window fits
        i64 groups = imin64(kv + 1, 1024 / (kl + 1));
        if (groups < 1) groups = 1;
        int threads = (int)imin64(1024, imax64(128, (((kl + 1) * groups + 31) / 32) * 32));
        gbtrf_window<<<1, threads, smem, h->stream>>>(m, n, (int)kl, (int)ku, dAB, ldab, d_ipiv, h->d_info, nslot);
        BMB_LAUNCH_CHECK(h);
    } else {
        rc = bmb_gbtrf_blocked(h, m, n, kl, ku, dAB, ldab, d_ipiv);
        if (rc) return rc;
    }
    int hinfo = 0;
    BMB_CUDA(h, cudaMemcpyAsync(&hinfo, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (info) *info = hinfo;
    return 0;
}
