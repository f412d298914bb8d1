// gbtrf_blocked.cu -- wide-band partial-pivot LU: right-looking panel + trailing update inside the 2kl+ku fill band.
//
// Same contract as gbtrf.cu (first-max pivots, reciprocal scaling, one FMA per element per eliminated column in
// ascending column order => pivots and factors bit-identical to DGBTF2; LAPACK's own blocked DGBTRF, which OpenBLAS
// runs for ku > 64, differs from that only by DGEMM rounding).
//
// For every panel of NB columns starting at J:
//   gbtrf_panel  (one CTA, panel of (NB+kl) x NB doubles in shared memory)
//       unblocked factorisation of the panel: block-wide IDAMAX, FULL-row swap inside the panel, reciprocal scaling,
//       rank-1 update of the remaining panel columns.  The multipliers of column j are written to AB right after
//       step j (LAPACK keeps them un-permuted); the fully swapped L panel goes to a workspace for the update.
//   gbtrf_update (one CTA per TC trailing columns, up to ju)
//       per column: the NB row interchanges, forward substitution with the unit-lower NB x NB block (rows of U),
//       then the Schur update x[i] = fma(-u[j], L[i,j], x[i]), j ascending, for the kl rows below (thread = row,
//       its L row held in registers, u broadcast from shared memory).
// `ju` (last column touched so far) and `info` live in device memory, so the host loop never synchronises.
#include "common.cuh"

#define PANEL_THREADS 1024

struct PanelState {  // h->d_info layout used by this file
    int info;        // [0]
    int pad;
    long long ju;    // [2..3] running 0-based ju
};

__global__ void gbtrf_zero_fill(i64 n, i64 kl, i64 kv, double *__restrict__ ab, i64 ldab, i64 ncols)
{
    // DGBTF2 zeroes rows [max(0,kv-c), kl) of every column c it can touch before using them as fill-in space
    const i64 total = kl * ncols;
    for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        const i64 c = t / kl, r = t - c * kl;
        if (r >= kv - c) ab[r + c * ldab] = 0.0;
    }
}

__global__ void __launch_bounds__(PANEL_THREADS, 1)
gbtrf_panel(i64 m, i64 n, int kl, int ku, double *__restrict__ ab, i64 ldab, i64 *__restrict__ ipiv, i64 J, int nbw,
            double *__restrict__ Lw, int PR, PanelState *__restrict__ st)
{
    extern __shared__ double P[];  // nbw columns x PR rows, P[c*PR + r] = A(J+r, J+c)
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ long long s_ju;
    __shared__ int s_info;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int kv = kl + ku;
    const i64 mn = m < n ? m : n;
    const int R = (int)(((i64)nbw + kl < m - J) ? (i64)nbw + kl : (m - J));  // panel rows J .. J+R-1
    if (tid == 0) { s_ju = st->ju; s_info = st->info; }
    // ---- load the panel (zero below the band: those slots are not stored) ----
    for (int c = 0; c < nbw; ++c) {
        const double *col = ab + (J + c) * ldab + (kv - c);
        for (int r = tid; r < R; r += PANEL_THREADS) P[c * PR + r] = (r <= c + kl && J + c < n) ? col[r] : 0.0;
    }
    __syncthreads();
    for (int jj = 0; jj < nbw; ++jj) {
        const i64 j = J + jj;
        if (j >= mn) break;
        const int km = (int)((kl < m - 1 - j) ? kl : (m - 1 - j));
        double *pc = P + jj * PR;
        // ---- IDAMAX over rows jj .. jj+km: first maximum ----
        double best = -1.0;
        int bidx = 0x7fffffff;
        for (int r = jj + tid; r <= jj + km; r += PANEL_THREADS) {
            const double v = fabs(pc[r]);
            if (v > best) { best = v; bidx = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bidx; }
        __syncthreads();
        if (wid == 0) {
            best = s_val[lane];
            bidx = s_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            if (lane == 0) {
                if (bidx > jj + km) bidx = jj;  // every candidate was NaN (v > best never held): keep the diagonal, stay in bounds
                s_piv = bidx;
                ipiv[j] = J + bidx + 1;
                if (pc[bidx] != 0.0) {
                    long long cand = j + ku + (bidx - jj);
                    if (cand > n - 1) cand = n - 1;
                    if (cand > s_ju) s_ju = cand;
                } else if (s_info == 0) {
                    s_info = (int)(j + 1);
                }
            }
        }
        __syncthreads();
        const int p = s_piv;
        const double pv = pc[p];
        if (pv != 0.0) {  // uniform
            if (p != jj && tid < nbw) {  // full-row swap inside the panel (keeps the workspace L fully permuted)
                const double t = P[tid * PR + jj];
                P[tid * PR + jj] = P[tid * PR + p];
                P[tid * PR + p] = t;
            }
            __syncthreads();
            const double rinv = 1.0 / pc[jj];
            double *gcol = ab + j * ldab + (kv - jj);  // AB(kv + r - jj, j)
            for (int r = jj + 1 + tid; r <= jj + km; r += PANEL_THREADS) {
                const double l = __dmul_rn(pc[r], rinv);
                pc[r] = l;
                gcol[r] = l;  // LAPACK format: the multiplier stays at the row it has NOW (later swaps do not move it)
                for (int c = jj + 1; c < nbw; ++c) P[c * PR + r] = fma(-P[c * PR + jj], l, P[c * PR + r]);
            }
            __syncthreads();
        } else {
            // exactly-zero pivot: DGBTF2 leaves the (all-zero) column alone; make AB agree with the panel copy
            double *gcol = ab + j * ldab + (kv - jj);
            for (int r = jj + 1 + tid; r <= jj + km; r += PANEL_THREADS) gcol[r] = pc[r];
        }
    }
    // ---- U11 (upper triangle of the panel) back to AB; fully swapped panel to the workspace ----
    for (int c = 0; c < nbw; ++c) {
        if (J + c >= n) break;
        double *col = ab + (J + c) * ldab + (kv - c);
        for (int r = tid; r <= c && r < R; r += PANEL_THREADS) col[r] = P[c * PR + r];
        for (int r = tid; r < R; r += PANEL_THREADS) Lw[(i64)c * PR + r] = P[c * PR + r];
    }
    if (tid == 0) { st->ju = s_ju; st->info = s_info; }
}

template <int NB>
__global__ void __launch_bounds__(PANEL_THREADS, 1)
gbtrf_update(i64 m, i64 n, int kl, int ku, double *__restrict__ ab, i64 ldab, const i64 *__restrict__ ipiv, i64 J,
             int nbw, const double *__restrict__ Lw, int PR, const PanelState *__restrict__ st, int TC)
{
    extern __shared__ double X[];  // TC columns x PR rows
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int kv = kl + ku;
    const i64 ju = st->ju;
    const i64 cbase = J + nbw + (i64)blockIdx.x * TC;
    if (cbase > ju || cbase >= n) return;
    const int nc = (int)((cbase + TC - 1 <= ju ? TC : (ju - cbase + 1)));
    const int R = (int)(((i64)nbw + kl < m - J) ? (i64)nbw + kl : (m - J));
    const i64 mn = m < n ? m : n;
    const int npiv = (int)((J + nbw <= mn) ? nbw : (mn - J));  // pivots actually taken in this panel
    // ---- load the column segments (rows J .. J+R-1); rows above the stored band are structural zeros ----
    for (int q = 0; q < nc; ++q) {
        const i64 c = cbase + q;
        const double *col = ab + c * ldab + (kv - (c - J));  // AB(kv + (J+r) - c, c) = col[r]
        const i64 rmin = c - kv - J;                         // first stored row, relative to J (may be <= 0)
        for (int r = tid; r < R; r += PANEL_THREADS) X[q * PR + r] = (r >= rmin && c < n) ? col[r] : 0.0;
    }
    __syncthreads();
    // ---- row interchanges, in pivot order (one thread per column) ----
    if (tid < nc) {
        double *x = X + tid * PR;
        for (int jj = 0; jj < npiv; ++jj) {
            const int p = (int)(ipiv[J + jj] - 1 - J);
            // rows above the stored band of this column are structural zeros DGBTF2 never swaps (c > j+kv >= ju(j))
            if (p != jj && jj >= cbase + tid - kv - J) { const double t = x[jj]; x[jj] = x[p]; x[p] = t; }
        }
    }
    __syncthreads();
    // ---- rows of U: forward substitution with the unit-lower npiv x npiv block (one warp per column) ----
    for (int q = wid; q < nc; q += PANEL_THREADS / 32) {
        double *x = X + q * PR;
        for (int jj = 0; jj < npiv; ++jj) {
            const double u = x[jj];
            __syncwarp();
            for (int i = jj + 1 + lane; i < npiv; i += 32) x[i] = fma(-u, Lw[(i64)jj * PR + i], x[i]);
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- Schur update of the rows below the pivot block: thread = row, L row in registers ----
    for (int r = npiv + tid; r < R; r += PANEL_THREADS) {
        double L[NB];
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) L[jj] = (jj < npiv) ? Lw[(i64)jj * PR + r] : 0.0;
        for (int q = 0; q < nc; ++q) {
            const double *x = X + q * PR;
            double acc = x[r];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj)
                if (jj < npiv) acc = fma(-x[jj], L[jj], acc);
            X[q * PR + r] = acc;
        }
    }
    __syncthreads();
    // ---- store back the stored part ----
    for (int q = 0; q < nc; ++q) {
        const i64 c = cbase + q;
        double *col = ab + c * ldab + (kv - (c - J));
        const i64 rmin = c - kv - J;
        for (int r = tid; r < R; r += PANEL_THREADS)
            if (r >= rmin && c < n) col[r] = X[q * PR + r];
    }
}

int bmb_gbtrf_pipe(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv, i64 *Jdone);
int bmb_gbtrf_strip(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv, i64 *Jdone);  // gbtrf_strip.cu

int bmb_gbtrf_blocked(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv)
{
    const i64 kv = kl + ku, mn = imin64(m, n);
    // panel width: the (NB+kl) x NB panel must fit in shared memory
    int NB = 16;
    while (NB > 2 && (size_t)(NB + kl + 1) * NB * sizeof(double) > 200 * 1024) NB >>= 1;
    int PR = (int)(NB + kl);
    PR |= 1;  // odd pitch: column-strided accesses of a row fall in different banks
    const size_t smem_p = (size_t)NB * PR * sizeof(double);
    if (smem_p > 220 * 1024) {
        snprintf(h->err, sizeof(h->err), "dgbtrf: kl = %lld is too wide for the blocked kernel", (long long)kl);
        return BMB200_ERR_CUDA;
    }
    int TC = (int)imin64(8, (200 * 1024) / ((size_t)PR * sizeof(double)));
    if (TC < 1) TC = 1;
    const size_t smem_u = (size_t)TC * PR * sizeof(double);
    int rc = bmb_ensure_scratch(h, (size_t)NB * PR * sizeof(double) + 512);
    if (rc) return rc;
    PanelState *st = (PanelState *)h->d_info;
    BMB_CUDA(h, cudaMemsetAsync(st, 0, sizeof(PanelState), h->stream));
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrf_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrf_update<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u));
    // fill-in rows of every column the factorisation can touch
    const i64 ncols = imin64(n, mn + kv);
    if (kl > 0) {
        const i64 total = kl * ncols;
        const int blocks = (int)imin64(cdiv64(total, 256), (i64)h->sm_count * 16);
        gbtrf_zero_fill<<<blocks, 256, 0, h->stream>>>(n, kl, kv, dAB, ldab, ncols);
        BMB_LAUNCH_CHECK(h);
    }
    // the bulk of the panels runs in the persistent pipelined kernel (gbtrf_pipe.cu); the stepwise kernels below
    // finish the last kl/NB + 1 panels (ragged rows) or do everything when the shape is not eligible
    // Interchange-free matrices (verified on the fly) take the strip-resident kernel (gbtrf_strip.cu); it hands back 0
    // columns, with AB as it was, when the shape is not eligible or an interchange turns out to be needed.
    i64 Jstart = 0;
    rc = bmb_gbtrf_strip(h, m, n, kl, ku, dAB, ldab, d_ipiv, &Jstart);
    if (rc) return rc;
    if (Jstart == 0) {
        rc = bmb_gbtrf_pipe(h, m, n, kl, ku, dAB, ldab, d_ipiv, &Jstart);
        if (rc) return rc;
    }
    rc = bmb_ensure_scratch(h, (size_t)NB * PR * sizeof(double) + 512);
    if (rc) return rc;
    double *Lw = (double *)h->scratch;  // (taken here: the kernels above may have grown, i.e. moved, the scratch buffer)
    const unsigned ublocks = (unsigned)cdiv64(kv, TC);
    for (i64 J = Jstart; J < mn; J += NB) {
        const int nbw = (int)imin64(NB, n - J);
        gbtrf_panel<<<1, PANEL_THREADS, smem_p, h->stream>>>(m, n, (int)kl, (int)ku, dAB, ldab, d_ipiv, J, nbw, Lw, PR, st);
        h->launches++;
        if (J + nbw < n && ublocks > 0) {
            gbtrf_update<16><<<ublocks, PANEL_THREADS, smem_u, h->stream>>>(m, n, (int)kl, (int)ku, dAB, ldab, d_ipiv, J,
                                                                           nbw, Lw, PR, st, TC);
            h->launches++;
        }
    }
    BMB_CUDA(h, cudaGetLastError());
    return 0;
}
