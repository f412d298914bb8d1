// gbmm.cu -- banded x banded and banded x dense products on sm_100a.
//
// bmb200_dgbmm_bb replaces _gbmm! (src/banded/gbmm.jl:296-340): the reference issues ONE dgbmv_ per
// column of C (m BLAS calls, three regimes + a trailing beta fill); here one launch computes
//     C[k,j] = beta*C[k,j] (or 0) then, for nu ascending over band(A row k) ^ band(B col j),
//              C[k,j] = fma(alpha*B[nu,j], A[k,nu], C[k,j])
// which is exactly the per-element operation order of that dgbmv_ sequence, so the
// result is bit-identical to the reference CPU path.
// bmb200_dgbmm_bd replaces the per-column mul! loop of src/generic/matmul.jl:243-256.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// v1 banded x banded: warp = output column, lanes = rows of that column (contiguous in C and in
// every A column), sweep over nu.  A is re-read from L1/L2 (each A column serves Bl+Bu+1 output
// columns).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gbmm_bb_sweep(i64 n, i64 nu, i64 m, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu, double alpha,
              const double *__restrict__ a, i64 lda, const double *__restrict__ b, i64 ldb, double beta,
              double *__restrict__ c, i64 ldc, i64 jbeg)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = jbeg + warp; j < m; j += nwarps) {
        i64 k0 = j - Cu; if (k0 < 0) k0 = 0;
        i64 k1 = j + Cl; if (k1 > n - 1) k1 = n - 1;
        // columns right of every B column that meets A (j >= nu+Bu): the reference beta-fills the WHOLE band
        // column, out-of-matrix slots included (_fill_lmul!(beta, view(C_data,:,nu+Bu+1:min(m,n+Cu))), gbmm.jl:339)
        if (j >= nu + Bu && j < n + Cu) { k0 = j - Cu; k1 = j + Cl; }
        i64 v0 = j - Bu; if (v0 < 0) v0 = 0;
        i64 v1 = j + Bl; if (v1 > nu - 1) v1 = nu - 1;
        const double *bcol = b + j * ldb + (Bu - j);   // B[v,j] = bcol[v]
        double *ccol = c + j * ldc + (Cu - j);         // C[k,j] = ccol[k]
        for (i64 kb = k0; kb <= k1; kb += 32) {
            const i64 k = kb + lane;
            const bool live = k <= k1;
            double acc = (beta == 0.0 || !live) ? 0.0 : __dmul_rn(beta, ccol[k]);
            if (alpha != 0.0) {
                // only nu with band(A col nu) touching rows [kb, kb+31]
                i64 w0 = kb - Al; if (w0 < v0) w0 = v0;
                i64 w1 = kb + 31 + Au; if (w1 > v1) w1 = v1;
#pragma unroll 4
                for (i64 v = w0; v <= w1; ++v) {
                    const double t = __dmul_rn(alpha, bcol[v]);
                    if (live && k >= v - Au && k <= v + Al) acc = fma(t, a[(Au + k - v) + v * lda], acc);
                }
            }
            if (live) ccol[k] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// gbmm_bb_dmma: banded x banded on the FP64 tensor cores (DMMA.8x8x4), for band widths whose tiles are dense
// enough.  In band storage every in-band tile is column-major dense with pitch lda-1, so C[K,J] += A[K,V] * B[V,J]
// over 8x8 (K,J) tiles and 4-wide V steps.  A CTA owns TJ output columns: the A columns [j0-Bu, j0+TJ+Bl) and the
// (alpha-scaled) B columns are staged once in shared memory with zero pads above/below each column, so out-of-band
// fragment elements read as 0.0 with no per-element test.  Column pitches are = 5 (mod 16) doubles, which makes
// both fragment loads (A: lane -> (row l/4, k l%4); B: lane -> (k l%4, col l/4)) bank-conflict free.
// DMMA.8x8x4 was verified on this GPU to equal the sequential FMA chain over k (tools/fp64_peaks.cu), and V is
// walked in ascending order from beta*C, so the result is bit-identical to the reference's per-column dgbmv_ chain.
// ------------------------------------------------------------------------------------------------
#define GM_TJ 32          // output columns per CTA
#define GM_PAD 11         // zero pad (doubles) above and below every staged column
#define GM_THREADS 256

int bmb_gbmm_wide(bmb200_ctx *h, i64 n, i64 nu, i64 mprod, i64 Al, i64 Au, i64 Bl, i64 Bu, i64 Cl, i64 Cu, double alpha, const double *dA,
                  i64 lda, const double *dB, i64 ldb, double beta, double *dC, i64 ldc);  // gbmm_wide.cu

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}
__host__ __device__ __forceinline__ i64 floordiv(i64 a, i64 b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// 8-byte asynchronous global->shared copy (LDGSTS); `valid == false` writes 0.0 without reading `src`
__device__ __forceinline__ void cp_async8_zfill(double *dst, const double *src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Flat staging of `ncols` whole band columns that are contiguous in global memory (ld == W) into the padded shared layout
// (pitch P): the slab is walked element by element by the whole CTA with an incrementally maintained (column, row), ~10 issued
// instructions per 32 elements -- the per-column loop (64-bit column address, row-range tests, a 3rd trip for row 64 alone) cost
// ~85 per 65-element column, and warp-stall sampling put 35-40 % of this kernel's time there.
__device__ __forceinline__ void gm_stage_flat(double *dst0, int P, const double *src, int W, int ncols, int tid, int nthreads)
{
    const int total = ncols * W;
    int col = tid / W, row = tid - col * W;
    const int dc = nthreads / W, dr = nthreads - dc * W;
    unsigned d = (unsigned)__cvta_generic_to_shared(dst0) + 8u * (unsigned)(col * P + row);
    const unsigned dstep = 8u * (unsigned)(dc * P + dr), dwrap = 8u * (unsigned)(P - W);
    const double *sp = src + tid;
    for (int e = tid; e < total; e += nthreads) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(sp) : "memory");
        sp += nthreads;
        d += dstep;
        row += dr;
        if (row >= W) { row -= W; d += dwrap; }
    }
}

// GM_NT = row tiles (of 8) per work item: the V range of an item is the union over its tiles, so a tall item spends
// most of its predicated DMMA slots outside the band parallelogram (9 tiles: 44 % of the slots live; 3 tiles: 77 %).
template <int GM_NT>
__global__ void __launch_bounds__(GM_THREADS, 2)
gbmm_bb_dmma(i64 n, i64 nu, i64 mcols, int Al, int Au, int Bl, int Bu, int Cl, int Cu, double alpha,
             const double *__restrict__ a, i64 lda, const double *__restrict__ b, i64 ldb, double beta,
             double *__restrict__ c, i64 ldc, int PA, int PB, int NA, i64 ntiles)
{
    extern __shared__ double sm[];
    double *As = sm;                       // NA columns x PA
    double *Bs = sm + (size_t)NA * PA;     // GM_TJ columns x PB
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int WA = Al + Au + 1, WB = Bl + Bu + 1;
    // zero everything once: the pads are never written again
    for (int t = tid; t < NA * PA + GM_TJ * PB; t += GM_THREADS) sm[t] = 0.0;
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;   // fragment coordinates: A (row fr, k fk), B (k fk, col fr)
    const int ntile_rows = (Cu + Cl + 8 + 7 + 7) / 8;          // row tiles that can touch one column block
    const int nchunks = (ntile_rows + GM_NT - 1) / GM_NT;
    const unsigned span = (unsigned)(Al + Au + 10);
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 j0 = tile * GM_TJ;
        const i64 vbase = 4 * floordiv(j0 - Bu, 4);  // first staged A column (aligned to the k step)
        // ---- stage A columns [vbase, vbase+NA) and B columns [j0, j0+TJ): asynchronous copies, all in flight ----
        // interior tiles (every staged column complete and inside the matrix, columns contiguous): flat slab copy
        const bool flatA = lda == WA && vbase >= Au && vbase + NA <= nu && vbase + NA - 1 + Al <= n - 1;
        const bool flatB = ldb == WB && j0 >= Bu && j0 + GM_TJ <= mcols && j0 + GM_TJ - 1 + Bl <= nu - 1;
        if (flatA) gm_stage_flat(As + GM_PAD, PA, a + vbase * lda, WA, NA, tid, GM_THREADS);
        if (flatB) gm_stage_flat(Bs + GM_PAD, PB, b + j0 * ldb, WB, GM_TJ, tid, GM_THREADS);
        for (int s = wid; s < (flatA ? 0 : NA); s += GM_THREADS / 32) {
            const i64 v = vbase + s;
            double *dst = As + (size_t)s * PA + GM_PAD;
            int rlo = 1, rhi = 0;  // band rows r with 0 <= v - Au + r < n
            if (v >= 0 && v < nu) {
                rlo = (v < Au) ? (int)(Au - v) : 0;
                rhi = (int)imin64_d((i64)WA - 1, n - 1 - v + Au);
            }
            const double *src = a + (rlo <= rhi ? v * lda : 0);
            if (rlo == 0 && rhi == WA - 1) {  // whole column inside the matrix (all but the first/last tiles)
                for (int r = lane; r < WA; r += 32) cp_async8_zfill(dst + r, src + r, true);
            } else {
                for (int r = lane; r < WA; r += 32) {
                    const bool ok = r >= rlo && r <= rhi;
                    cp_async8_zfill(dst + r, ok ? src + r : a, ok);
                }
            }
        }
        for (int s = wid; s < (flatB ? 0 : GM_TJ); s += GM_THREADS / 32) {
            const i64 j = j0 + s;
            double *dst = Bs + (size_t)s * PB + GM_PAD;
            int rlo = 1, rhi = 0;  // band rows r with 0 <= j - Bu + r < nu
            if (j < mcols) {
                rlo = (j < Bu) ? (int)(Bu - j) : 0;
                rhi = (int)imin64_d((i64)WB - 1, nu - 1 - j + Bu);
            }
            const double *src = b + (rlo <= rhi ? j * ldb : 0);
            if (rlo == 0 && rhi == WB - 1) {
                for (int r = lane; r < WB; r += 32) cp_async8_zfill(dst + r, src + r, true);
            } else {
                for (int r = lane; r < WB; r += 32) {
                    const bool ok = r >= rlo && r <= rhi;
                    cp_async8_zfill(dst + r, ok ? src + r : b, ok);
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        // ---- work items: (column block of 8) x (chunk of GM_NT row tiles); everything relative to j0 in int ----
        const int kfirst_r = (int)(8 * floordiv(j0 - Cu, 8) - j0);     // first row tile of column block 0
        const int vbase_r = (int)(vbase - j0);
        for (int item = wid; item < (GM_TJ / 8) * nchunks; item += GM_THREADS / 32) {
            const int jb = item / nchunks, ch = item - jb * nchunks;
            const int jc0r = 8 * jb;
            if (j0 + jc0r >= mcols) continue;
            const int k0r = kfirst_r + 8 * jb + 8 * GM_NT * ch;        // first row of this chunk
            double acc[GM_NT][2];
            // lane owns C(k0 + 8t + fr, jc0 + 2*fk + {0,1}) = cp0[8t + e*(ldc-1)]; rows and columns of an interior item are
            // all inside the matrix, so only the band test remains (and none at all for a tile fully inside the band)
            const bool inner = j0 + jc0r + 7 < mcols && j0 + k0r >= 0 && j0 + k0r + 8 * GM_NT - 1 < n;
            double *cp0 = c + (Cu + (k0r + fr) - (jc0r + 2 * fk)) + (j0 + jc0r + 2 * fk) * ldc;
            if (beta == 0.0) {
#pragma unroll
                for (int t = 0; t < GM_NT; ++t) acc[t][0] = acc[t][1] = 0.0;
            } else {
#pragma unroll
                for (int t = 0; t < GM_NT; ++t) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kr = k0r + 8 * t + fr, jr = jc0r + 2 * fk + e;
                        const i64 k = j0 + kr, j = j0 + jr;
                        const bool in = j < mcols && k >= 0 && k < n && kr - jr <= Cl && jr - kr <= Cu;
                        acc[t][e] = in ? __dmul_rn(beta, cp0[8 * t + e * (ldc - 1)]) : 0.0;
                    }
                }
            }
            // V range: band of B over these columns, intersected with the band of A over these rows and [0, nu)
            int v0 = max(jc0r - Bu, k0r - Al), v1 = min(jc0r + 7 + Bl, k0r + 8 * GM_NT - 1 + Au);
            if ((i64)v0 < -j0) v0 = (int)(-j0);
            if ((i64)v1 > nu - 1 - j0) v1 = (int)(nu - 1 - j0);
            v0 = (int)(4 * floordiv(j0 + v0, 4) - j0);                 // aligned like vbase
            // B(v, j) -> Bs[jr*PB + PAD + v - jr + Bu];  A(k, v) -> As[(v - vbase)*PA + PAD + k - v + Au]
            const double *bp = Bs + (size_t)(jc0r + fr) * PB + GM_PAD + Bu - (jc0r + fr) + fk + v0;
            const double *ap = As + (size_t)(v0 + fk - vbase_r) * PA + GM_PAD + Au - (v0 + fk) + k0r + fr;
            int d0 = k0r - v0 + Au + 7;   // tile t meets the band of A columns [v, v+3] iff 0 <= d0 + 8t <= span
            const int astep = 4 * (PA - 1);
            for (int v = v0; v <= v1; v += 4) {
                const double bf = (alpha == 1.0) ? *bp : __dmul_rn(alpha, *bp);
#pragma unroll
                for (int t = 0; t < GM_NT; ++t) {
                    if ((unsigned)(d0 + 8 * t) <= span) {  // warp-uniform
                        const double af = ap[8 * t];
                        dmma884(acc[t][0], acc[t][1], af, bf);
                    }
                }
                bp += 4;
                ap += astep;
                d0 -= 4;
            }
            if (inner) {
#pragma unroll
                for (int t = 0; t < GM_NT; ++t) {
                    const int dlo = k0r + 8 * t - jc0r - 7, dhi = dlo + 14;  // row - column over the 8 x 8 tile
                    if (dhi <= Cl && dlo >= -Cu) {                            // warp-uniform: tile fully inside the band
                        cp0[8 * t] = acc[t][0];
                        cp0[8 * t + (ldc - 1)] = acc[t][1];
                    } else {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int dd = k0r + 8 * t + fr - (jc0r + 2 * fk + e);
                            if (dd <= Cl && -dd <= Cu) cp0[8 * t + e * (ldc - 1)] = acc[t][e];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < GM_NT; ++t) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kr = k0r + 8 * t + fr, jr = jc0r + 2 * fk + e;
                        const i64 k = j0 + kr, j = j0 + jr;
                        if (j < mcols && k >= 0 && k < n && kr - jr <= Cl && jr - kr <= Cu) cp0[8 * t + e * (ldc - 1)] = acc[t][e];
                    }
                }
            }
        }
        __syncthreads();  // the staged columns are overwritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------
// gbmm_bb_ring: the same tile arithmetic as gbmm_bb_dmma, restructured around the two things that bounded it (ncu,
// profiles/gbmm_c3_r1_ncu.md): every CTA tile re-staged NA = TJ+Bl+Bu+7 columns of A for TJ output columns (3.2x the
// algorithmic A traffic at C3, 31 % of all issued instructions) and nothing overlapped the staging.  Here a CTA walks
// CONSECUTIVE column tiles and keeps the staged A columns in a ring: per tile only the TJ new columns (and the next
// tile's B columns, double-buffered) are fetched, with cp.async, while the tensor cores work on the current tile.
// One persistent CTA per SM, 16 warps; per-element operation order unchanged => bit-identical results.
// ------------------------------------------------------------------------------------------------
template <int GM_NT, int GR_THREADS>
__global__ void __launch_bounds__(GR_THREADS, 1)
gbmm_bb_ring(i64 n, i64 nu, i64 mcols, int Al, int Au, int Bl, int Bu, int Cl, int Cu, double alpha,
             const double *__restrict__ a, i64 lda, const double *__restrict__ b, i64 ldb, double beta,
             double *__restrict__ c, i64 ldc, int PA, int PB, int NA, int RS, int TJ, i64 ntiles, i64 tiles_per_cta)
{
    extern __shared__ double sm[];
    double *As = sm;                             // RS ring slots x PA
    double *Bs0 = sm + (size_t)RS * PA;          // TJ columns x PB, two buffers
    double *Bs1 = Bs0 + (size_t)TJ * PB;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = GR_THREADS / 32;
    const int WA = Al + Au + 1, WB = Bl + Bu + 1;
    const i64 t_begin = blockIdx.x * tiles_per_cta;
    const i64 t_end = (t_begin + tiles_per_cta < ntiles) ? t_begin + tiles_per_cta : ntiles;
    if (t_begin >= t_end) return;
    for (int t = tid; t < RS * PA + 2 * TJ * PB; t += GR_THREADS) sm[t] = 0.0;  // pads are never written again
    __syncthreads();
    const i64 vorg = 4 * floordiv(t_begin * TJ - Bu, 4);  // ring origin: column v lives in slot (v - vorg) mod RS
    auto stageA = [&](i64 vfirst, int ncols) {
        if (ncols >= 16 && lda == WA && vfirst >= Au && vfirst + ncols <= nu && vfirst + ncols - 1 + Al <= n - 1) {  // interior: flat slab copy
            const int slot0 = (int)((unsigned)(vfirst - vorg) % (unsigned)RS);
            const int n1 = (RS - slot0 < ncols) ? RS - slot0 : ncols;                                  // columns before the ring wraps
            gm_stage_flat(As + (size_t)slot0 * PA + GM_PAD, PA, a + vfirst * lda, WA, n1, tid, GR_THREADS);
            if (n1 < ncols) gm_stage_flat(As + GM_PAD, PA, a + (vfirst + n1) * lda, WA, ncols - n1, tid, GR_THREADS);
            return;
        }
        for (int s = wid; s < ncols; s += NW) {
            const i64 v = vfirst + s;
            double *dst = As + (size_t)((unsigned)(v - vorg) % (unsigned)RS) * PA + GM_PAD;
            int rlo = 1, rhi = 0;  // band rows r with 0 <= v - Au + r < n
            if (v >= 0 && v < nu) {
                rlo = (v < Au) ? (int)(Au - v) : 0;
                rhi = (int)imin64_d((i64)WA - 1, n - 1 - v + Au);
            }
            const double *src = a + (rlo <= rhi ? v * lda : 0);
            if (rlo == 0 && rhi == WA - 1) {
                for (int r = lane; r < WA; r += 32) cp_async8_zfill(dst + r, src + r, true);
            } else {
                for (int r = lane; r < WA; r += 32) {
                    const bool ok = r >= rlo && r <= rhi;
                    cp_async8_zfill(dst + r, ok ? src + r : a, ok);
                }
            }
        }
    };
    auto stageB = [&](double *Bs, i64 j0) {
        if (TJ >= 16 && ldb == WB && j0 >= Bu && j0 + TJ <= mcols && j0 + TJ - 1 + Bl <= nu - 1) {
            gm_stage_flat(Bs + GM_PAD, PB, b + j0 * ldb, WB, TJ, tid, GR_THREADS);
            return;
        }
        for (int s = wid; s < TJ; s += NW) {
            const i64 j = j0 + s;
            double *dst = Bs + (size_t)s * PB + GM_PAD;
            int rlo = 1, rhi = 0;  // band rows r with 0 <= j - Bu + r < nu
            if (j < mcols) {
                rlo = (j < Bu) ? (int)(Bu - j) : 0;
                rhi = (int)imin64_d((i64)WB - 1, nu - 1 - j + Bu);
            }
            const double *src = b + (rlo <= rhi ? j * ldb : 0);
            if (rlo == 0 && rhi == WB - 1) {
                for (int r = lane; r < WB; r += 32) cp_async8_zfill(dst + r, src + r, true);
            } else {
                for (int r = lane; r < WB; r += 32) {
                    const bool ok = r >= rlo && r <= rhi;
                    cp_async8_zfill(dst + r, ok ? src + r : b, ok);
                }
            }
        }
    };
    stageA(vorg, NA);
    stageB(Bs0, t_begin * TJ);
    cp_async_wait_all();
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;   // fragment coordinates: A (row fr, k fk), B (k fk, col fr)
    const int ntile_rows = (Cu + Cl + 8 + 7 + 7) / 8;          // row tiles that can touch one 8-column block
    const int nchunks = (ntile_rows + GM_NT - 1) / GM_NT;
    const int items_per_ct = (TJ / 8) * nchunks;               // work items (8 columns x GM_NT row tiles) of one CTA tile
    const unsigned span = (unsigned)(Al + Au + 10);
    const int ringlen = RS * PA;
    const int astep = 4 * (PA - 1);
    const bool scale = alpha != 1.0;
    // Items differ in length (the V range shrinks towards the band edges), so warps draw them from a ticket counter
    // instead of a fixed round-robin: the per-tile barrier then waits for at most one item.
    __shared__ unsigned ticket_ctr;
    if (tid == 0) ticket_ctr = NW;  // tickets 0..NW-1 are handed out statically
    unsigned ticket = wid;          // next item of this warp, numbered across all CTA tiles of this CTA
    __syncthreads();
    for (i64 tile = t_begin; tile < t_end; ++tile) {
        const i64 j0 = tile * TJ;
        const i64 vbase = vorg + (tile - t_begin) * TJ;  // = 4*floor((j0-Bu)/4): TJ is a multiple of 4
        const int seq = (int)(tile - t_begin);
        const double *Bs = (seq & 1) ? Bs1 : Bs0;
        if (tile + 1 < t_end) {  // next tile's new A columns and its B columns, in flight during this tile's DMMAs
            stageA(vbase + NA, TJ);
            stageB((seq & 1) ? Bs0 : Bs1, j0 + TJ);
        }
        const int kfirst_r = (int)(8 * floordiv(j0 - Cu, 8) - j0);
        const int vbase_r = (int)(vbase - j0);
        const int slot_base = (int)((unsigned)(vbase - vorg) % (unsigned)RS);
        const unsigned tk0 = (unsigned)seq * (unsigned)items_per_ct, tk1 = tk0 + (unsigned)items_per_ct;
        while (ticket < tk1) {
            const int item = (int)(ticket - tk0);
            const int jb = item / nchunks, ch = item - jb * nchunks;
            const int jc0r = 8 * jb;
            if (j0 + jc0r < mcols) {
                const int k0r = kfirst_r + 8 * jb + 8 * GM_NT * ch;
                double acc[GM_NT][2];
                const bool inner = j0 + jc0r + 7 < mcols && j0 + k0r >= 0 && j0 + k0r + 8 * GM_NT - 1 < n;
                double *cp0 = c + (Cu + (k0r + fr) - (jc0r + 2 * fk)) + (j0 + jc0r + 2 * fk) * ldc;
                if (beta == 0.0) {
#pragma unroll
                    for (int t = 0; t < GM_NT; ++t) acc[t][0] = acc[t][1] = 0.0;
                } else {
#pragma unroll
                    for (int t = 0; t < GM_NT; ++t) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int kr = k0r + 8 * t + fr, jr = jc0r + 2 * fk + e;
                            const i64 k = j0 + kr, j = j0 + jr;
                            const bool in = j < mcols && k >= 0 && k < n && kr - jr <= Cl && jr - kr <= Cu;
                            acc[t][e] = in ? __dmul_rn(beta, cp0[8 * t + e * (ldc - 1)]) : 0.0;
                        }
                    }
                }
                int v0 = max(jc0r - Bu, k0r - Al), v1 = min(jc0r + 7 + Bl, k0r + 8 * GM_NT - 1 + Au);
                if ((i64)v0 < -j0) v0 = (int)(-j0);
                if ((i64)v1 > nu - 1 - j0) v1 = (int)(nu - 1 - j0);
                v0 = (int)(4 * floordiv(j0 + v0, 4) - j0);
                // B(v, j) -> Bs[jr*PB + PAD + v - jr + Bu];  A(k, v) -> As[slot(v)*PA + PAD + k - v + Au]
                const double *bp = Bs + (size_t)(jc0r + fr) * PB + GM_PAD + Bu - (jc0r + fr) + fk + v0;
                int aslot = slot_base + (v0 - vbase_r);  // ring slot of column v0: a multiple of 4, the ring wraps between steps
                if (aslot >= RS) aslot -= RS;
                const double *ap = As + (size_t)(aslot + fk) * PA + GM_PAD + Au - (v0 + fk) + k0r + fr;
                int d0 = k0r - v0 + Au + 7;   // tile t meets the band of A columns [v, v+3] iff 0 <= d0 + 8t <= span
                const int nsteps = (v1 >= v0) ? (v1 - v0) / 4 + 1 : 0;
                int n1 = (RS - aslot) / 4;    // steps before the ring wraps
                if (n1 > nsteps) n1 = nsteps;
                for (int part = 0; part < 2; ++part) {
                    const int ns = part ? nsteps - n1 : n1;
                    if (part) ap -= ringlen;
                    for (int st = 0; st < ns; ++st) {
                        double bf = *bp;
                        if (scale) bf = __dmul_rn(alpha, bf);
#pragma unroll
                        for (int t = 0; t < GM_NT; ++t) {
                            if ((unsigned)(d0 + 8 * t) <= span) {  // warp-uniform
                                const double af = ap[8 * t];
                                dmma884(acc[t][0], acc[t][1], af, bf);
                            }
                        }
                        bp += 4;
                        ap += astep;
                        d0 -= 4;
                    }
                }
                if (inner) {
#pragma unroll
                    for (int t = 0; t < GM_NT; ++t) {
                        const int dlo = k0r + 8 * t - jc0r - 7, dhi = dlo + 14;
                        if (dhi <= Cl && dlo >= -Cu) {
                            cp0[8 * t] = acc[t][0];
                            cp0[8 * t + (ldc - 1)] = acc[t][1];
                        } else {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int dd = k0r + 8 * t + fr - (jc0r + 2 * fk + e);
                                if (dd <= Cl && -dd <= Cu) cp0[8 * t + e * (ldc - 1)] = acc[t][e];
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < GM_NT; ++t) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int kr = k0r + 8 * t + fr, jr = jc0r + 2 * fk + e;
                            const i64 k = j0 + kr, j = j0 + jr;
                            if (j < mcols && k >= 0 && k < n && kr - jr <= Cl && jr - kr <= Cu) cp0[8 * t + e * (ldc - 1)] = acc[t][e];
                        }
                    }
                }
            }
            unsigned nt = 0;
            if (lane == 0) nt = atomicAdd(&ticket_ctr, 1u);
            ticket = __shfl_sync(0xffffffffu, nt, 0);
        }
        cp_async_wait_all();
        __syncthreads();  // next tile's columns have landed; this tile's oldest TJ ring slots may be overwritten
    }
}

static int pitch5(int need)  // smallest p >= need with p = 5 (mod 16)
{
    int p = need;
    while ((p & 15) != 5) ++p;
    return p;
}

extern "C" int bmb200_dgbmm_bb(bmb200_handle_t h, int64_t n, int64_t nu, int64_t m, int64_t Al, int64_t Au,
                               int64_t Bl, int64_t Bu, int64_t Cl, int64_t Cu, double alpha, const double *dA,
                               int64_t lda, const double *dB, int64_t ldb, double beta, double *dC, int64_t ldc)
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
    if (lda < Al + Au + 1) return -13;
    if (ldb < Bl + Bu + 1) return -15;
    if (ldc < Cl + Cu + 1) return -18;
    if (n == 0 || m == 0) return 0;
    if (!dC || (nu > 0 && (!dA || !dB))) return -12;
    DeviceGuard g(h->device);
    const int threads = 256;
    // columns j >= nu+Bu only get the beta-fill (whole band columns, like gbmm.jl:339): sweep kernel.
    // The product columns go to the tensor-core kernel when the band is wide enough for dense tiles and the
    // staged tiles fit in shared memory; otherwise to the sweep kernel as well.
    i64 jsplit = 0;
    h->last_gbmm_path = 0;
    const i64 mprod = imin64(m, nu + Bu);
    const i64 WA = Al + Au + 1, WB = Bl + Bu + 1;
    if (alpha != 0.0 && mprod > 0 && WA >= 9 && WB >= 9 && Cl == imin64(n - 1, Al + Bl) && Cu == imin64(m - 1, Au + Bu)) {
        const int PA = pitch5((int)WA + 2 * GM_PAD), PB = pitch5((int)WB + 2 * GM_PAD);
        // Two tensor-core kernels: the two-CTA-per-SM tile kernel while a tile's staged columns fit in half an SM's shared
        // memory (C3: 4.45 ms), else the persistent ring kernel (one CTA per SM; C3: 4.63 ms, (64,64)x(64,64): 12.4 ms
        // where the sweep kernel took 110 ms).  tune.gbmm_ring = 1 forces the ring kernel, 0 disables it.
        const int ring_env = h->tune.gbmm_ring;
        if (h->tune.gbmm_wide == 1 && bmb_gbmm_wide(h, n, nu, mprod, Al, Au, Bl, Bu, Cl, Cu, alpha, dA, lda, dB, ldb, beta, dC, ldc) == 0) {
            jsplit = mprod;
            h->last_gbmm_path = 3;
        }
        const bool tile_fits = ((size_t)(GM_TJ + Bl + Bu + 7) * PA + (size_t)GM_TJ * PB) * sizeof(double) <= 110 * 1024;
        const bool try_ring = jsplit == 0 && (ring_env == 1 || (ring_env != 0 && !tile_fits));
        for (int TJ = 32; TJ >= 8 && try_ring && jsplit == 0; TJ >>= 1) {  // ring kernel: widest tile whose ring fits
            const int NAr = (int)(TJ + Bl + Bu + 4 + 3);
            const int RS = 4 * ((NAr + TJ + 3) / 4);
            const size_t smem_r = ((size_t)RS * PA + 2 * (size_t)TJ * PB) * sizeof(double);
            if (smem_r > 225 * 1024) continue;
            const i64 ntiles = cdiv64(mprod, TJ);
            const i64 blocks = imin64(ntiles, (i64)h->sm_count);
            const i64 tpc = cdiv64(ntiles, blocks);
            if (tpc * TJ + Bl + Bu + 64 >= ((i64)1 << 31)) break;
            const int rnt = h->tune.gbmm_nt;
#define GR_LAUNCH(NT, TH)                                                                                                              \
    do {                                                                                                                               \
        BMB_CUDA(h, cudaFuncSetAttribute(gbmm_bb_ring<NT, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));             \
        gbmm_bb_ring<NT, TH><<<(unsigned)cdiv64(ntiles, tpc), TH, smem_r, h->stream>>>(n, nu, mprod, (int)Al, (int)Au, (int)Bl,        \
                                                                                       (int)Bu, (int)Cl, (int)Cu, alpha, dA, lda,     \
                                                                                       dB, ldb, beta, dC, ldc, PA, PB, NAr, RS,       \
                                                                                       TJ, ntiles, tpc);                              \
    } while (0)
            const int rw_env = h->tune.gbmm_rw;  // warps per CTA
            const int rw = rw_env ? rw_env : (TJ >= 16 ? 16 : 12);
            if (rw == 12) { if (rnt == 2) GR_LAUNCH(2, 384); else GR_LAUNCH(3, 384); }
            else { if (rnt == 2) GR_LAUNCH(2, 512); else GR_LAUNCH(3, 512); }
#undef GR_LAUNCH
            BMB_LAUNCH_CHECK(h);
            jsplit = mprod;
            h->last_gbmm_path = 2;
        }
        const int NA = (int)(GM_TJ + Bl + Bu + 4 + 3);
        const size_t smem = ((size_t)NA * PA + (size_t)GM_TJ * PB) * sizeof(double);
        if (jsplit == 0 && smem <= 110 * 1024) {
            const i64 ntiles = cdiv64(mprod, GM_TJ);
            const i64 blocks = imin64(ntiles, (i64)h->sm_count * 2);
            const int nt_env = h->tune.gbmm_nt;
#define GM_LAUNCH(NT)                                                                                                          \
    do {                                                                                                                       \
        BMB_CUDA(h, cudaFuncSetAttribute(gbmm_bb_dmma<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        gbmm_bb_dmma<NT><<<(unsigned)blocks, GM_THREADS, smem, h->stream>>>(n, nu, mprod, (int)Al, (int)Au, (int)Bl, (int)Bu,  \
                                                                            (int)Cl, (int)Cu, alpha, dA, lda, dB, ldb, beta,  \
                                                                            dC, ldc, PA, PB, NA, ntiles);                     \
    } while (0)
            switch (nt_env) {
            case 2: GM_LAUNCH(2); break;
            case 4: GM_LAUNCH(4); break;
            case 6: GM_LAUNCH(6); break;
            case 9: GM_LAUNCH(9); break;
            default: GM_LAUNCH(3); break;
            }
#undef GM_LAUNCH
            BMB_LAUNCH_CHECK(h);
            jsplit = mprod;
            h->last_gbmm_path = 1;
        }
        // bands too wide for staged whole columns: the K-blocked tensor-core kernel (gbmm_wide.cu)
        if (jsplit == 0 && h->tune.gbmm_wide != 0 &&
            bmb_gbmm_wide(h, n, nu, mprod, Al, Au, Bl, Bu, Cl, Cu, alpha, dA, lda, dB, ldb, beta, dC, ldc) == 0) {
            jsplit = mprod;
            h->last_gbmm_path = 3;
        }
    }
    if (jsplit < m) {
        const i64 blocks = imin64(cdiv64(m - jsplit, threads / 32), (i64)h->sm_count * 8);
        gbmm_bb_sweep<<<(unsigned)blocks, threads, 0, h->stream>>>(n, nu, m, Al, Au, Bl, Bu, Cl, Cu, alpha, dA, lda, dB,
                                                                   ldb, beta, dC, ldc, jsplit);
        BMB_LAUNCH_CHECK(h);
    }
    return 0;
}

extern "C" int bmb200_internal_last_gbmm_path(bmb200_handle_t h) { return h ? h->last_gbmm_path : -1; }

// ------------------------------------------------------------------------------------------------
// banded x dense ('N'): lane = row, NR right-hand sides per thread share every A load.
// ------------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(256)
gbmm_bd_n(i64 m, i64 n, i64 kl, i64 ku, i64 nrhs, double alpha, const double *__restrict__ a, i64 lda,
          const double *__restrict__ b, i64 ldb, double beta, double *__restrict__ c, i64 ldc, i64 row_tiles)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const i64 rhs_tiles = (nrhs + NR - 1) / NR;
    const i64 total = row_tiles * rhs_tiles;
    for (i64 w = warp; w < total; w += nwarps) {
        const i64 rt = w % row_tiles, ct = w / row_tiles;
        const i64 i = rt * 32 + lane, c0 = ct * NR;
        const bool live = i < m;
        double acc[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q)
            acc[q] = (beta == 0.0 || !live || c0 + q >= nrhs) ? 0.0 : __dmul_rn(beta, c[i + (c0 + q) * ldc]);
        i64 jlo = rt * 32 - kl; if (jlo < 0) jlo = 0;
        i64 jhi = rt * 32 + 31 + ku; if (jhi > n - 1) jhi = n - 1;
        if (alpha != 0.0)
            for (i64 j = jlo; j <= jhi; ++j) {
                const bool in = live && i >= j - ku && i <= j + kl;
                const double av = in ? a[(ku + i - j) + j * lda] : 0.0;
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    if (c0 + q < nrhs) {
                        const double t = __dmul_rn(alpha, b[j + (c0 + q) * ldb]);
                        if (in) acc[q] = fma(t, av, acc[q]);
                    }
                }
            }
#pragma unroll
        for (int q = 0; q < NR; ++q)
            if (live && c0 + q < nrhs) c[i + (c0 + q) * ldc] = acc[q];
    }
}

// banded^T x dense: warp = output row j (column of A), shuffle reduction per right-hand side.
template <int NR>
__global__ void __launch_bounds__(256)
gbmm_bd_t(i64 m, i64 n, i64 kl, i64 ku, i64 nrhs, double alpha, const double *__restrict__ a, i64 lda,
          const double *__restrict__ b, i64 ldb, double beta, double *__restrict__ c, i64 ldc)
{
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const i64 rhs_tiles = (nrhs + NR - 1) / NR;
    const i64 total = n * rhs_tiles;
    for (i64 w = warp; w < total; w += nwarps) {
        const i64 j = w % n, c0 = (w / n) * NR;
        i64 i0 = j - ku; if (i0 < 0) i0 = 0;
        i64 i1 = j + kl; if (i1 > m - 1) i1 = m - 1;
        const double *colp = a + j * lda + (ku - j);
        double temp[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) temp[q] = 0.0;
        for (i64 i = i0 + lane; i <= i1; i += 32) {
            const double av = colp[i];
#pragma unroll
            for (int q = 0; q < NR; ++q)
                if (c0 + q < nrhs) temp[q] = fma(av, b[i + (c0 + q) * ldb], temp[q]);
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) temp[q] += __shfl_xor_sync(0xffffffffu, temp[q], o);
            if (lane == 0 && c0 + q < nrhs) {
                double *p = c + j + (c0 + q) * ldc;
                const double y0 = (beta == 0.0) ? 0.0 : __dmul_rn(beta, *p);
                *p = fma(alpha, temp[q], y0);
            }
        }
    }
}

extern "C" int bmb200_dgbmm_bd(bmb200_handle_t h, char trans, int64_t m, int64_t n, int64_t kl, int64_t ku,
                               int64_t nrhs, double alpha, const double *dA, int64_t lda, const double *dB,
                               int64_t ldb, double beta, double *dC, int64_t ldc)
{
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (kl < 0) return -5;
    if (ku < 0) return -6;
    if (nrhs < 0) return -7;
    if (lda < kl + ku + 1) return -10;
    const i64 rowsB = tr ? m : n, rowsC = tr ? n : m;
    if (ldb < imax64(1, rowsB)) return -12;
    if (ldc < imax64(1, rowsC)) return -15;
    if (rowsC == 0 || nrhs == 0) return 0;
    DeviceGuard g(h->device);
    if (rowsB == 0 || alpha == 0.0) return bmb200_dfill_lmul(h, beta, dC, rowsC, nrhs, ldc, 1);
    const i64 kle = imin64(kl, m - 1), kue = imin64(ku, n - 1);
    const double *ae = dA + (ku - kue);
    const int threads = 256;
    constexpr int NR = 8;
    if (!tr) {
        const i64 row_tiles = cdiv64(m, 32);
        const i64 total = row_tiles * cdiv64(nrhs, NR);
        const i64 blocks = imin64(cdiv64(total, threads / 32), (i64)h->sm_count * 8);
        gbmm_bd_n<NR><<<(unsigned)blocks, threads, 0, h->stream>>>(m, n, kle, kue, nrhs, alpha, ae, lda, dB, ldb, beta,
                                                                   dC, ldc, row_tiles);
    } else {
        const i64 total = n * cdiv64(nrhs, NR);
        const i64 blocks = imin64(cdiv64(total, threads / 32), (i64)h->sm_count * 8);
        gbmm_bd_t<NR><<<(unsigned)blocks, threads, 0, h->stream>>>(m, n, kle, kue, nrhs, alpha, ae, lda, dB, ldb, beta,
                                                                   dC, ldc);
    }
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// dense x banded: C <- alpha * A * op(P) + beta * C, A dense M x K, C dense M x N (both column-major).
// Replaces the per-ROW loop of src/generic/matmul.jl:258-271 (one strided gbmv per row of C: M BLAS calls) by one launch.
//   trans == 'N': op(P) = P, K x N with (kl,ku): row i of C is gbmv('T', P, A[i,:]) -- a dot product per entry (the order of
//                 OpenBLAS' dgbmv_t dot is unspecified; same formula as gbmv_t_* here: temp = sum fma, C = fma(alpha, temp, beta*C)).
//   trans == 'T': op(P) = P^T, P is N x K with (kl,ku): row i of C is gbmv('N', P, A[i,:]) -- every C[i,j] accumulates over k
//                 ascending with t = alpha*A[i,k] rounded first, one FMA per term: bit-identical to the reference's dgbmv_n rows.
// thread = one entry of C, i fastest: A and C accesses are coalesced down the columns, the band entry is a warp broadcast.
// ------------------------------------------------------------------------------------------------
template <bool TR>
__global__ void __launch_bounds__(256)
gbmm_db_kernel(i64 M, i64 K, i64 N, i64 kl, i64 ku, double alpha, const double *__restrict__ a, i64 lda,
               const double *__restrict__ p, i64 ldp, double beta, double *__restrict__ c, i64 ldc)
{
    const i64 total = M * N;
    for (i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        const i64 j = t / M, i = t - j * M;
        double *cp = c + i + j * ldc;
        const double c0 = (beta == 0.0) ? 0.0 : __dmul_rn(beta, *cp);
        if (!TR) {
            i64 k0 = j - ku; if (k0 < 0) k0 = 0;
            i64 k1 = j + kl; if (k1 > K - 1) k1 = K - 1;
            const double *col = p + j * ldp + (ku - j);  // P[k,j] = col[k]
            double temp = 0.0;
            for (i64 k = k0; k <= k1; ++k) temp = fma(col[k], a[i + k * lda], temp);
            *cp = fma(alpha, temp, c0);
        } else {
            i64 k0 = j - kl; if (k0 < 0) k0 = 0;
            i64 k1 = j + ku; if (k1 > K - 1) k1 = K - 1;
            double acc = c0;
            for (i64 k = k0; k <= k1; ++k)  // P[j,k] = p[(ku + j - k) + k*ldp]
                acc = fma(__dmul_rn(alpha, a[i + k * lda]), p[(ku + j - k) + k * ldp], acc);
            *cp = acc;
        }
    }
}

extern "C" int bmb200_dgbmm_db(bmb200_handle_t h, char trans, int64_t M, int64_t K, int64_t N, int64_t kl, int64_t ku,
                               double alpha, const double *dA, int64_t lda, const double *dP, int64_t ldp, double beta,
                               double *dC, int64_t ldc)
{
    if (!h) return -1;
    const bool tr = (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c');
    if (!tr && !(trans == 'N' || trans == 'n')) return -2;
    if (M < 0) return -3;
    if (K < 0) return -4;
    if (N < 0) return -5;
    if (kl < 0) return -6;
    if (ku < 0) return -7;
    if (lda < imax64(1, M)) return -10;
    if (ldp < kl + ku + 1) return -12;
    if (ldc < imax64(1, M)) return -15;
    if (M == 0 || N == 0) return 0;
    DeviceGuard g(h->device);
    const i64 blocks = imin64(cdiv64(M * N, 256), (i64)h->sm_count * 16);
    if (tr) gbmm_db_kernel<true><<<(unsigned)blocks, 256, 0, h->stream>>>(M, K, N, kl, ku, alpha, dA, lda, dP, ldp, beta, dC, ldc);
    else gbmm_db_kernel<false><<<(unsigned)blocks, 256, 0, h->stream>>>(M, K, N, kl, ku, alpha, dA, lda, dP, ldp, beta, dC, ldc);
    BMB_LAUNCH_CHECK(h);
    return 0;
}
