// pb.cu -- banded Cholesky on the device: pbtrf! / pbtrs! of the reference (src/lapack.jl:268-332), reached from
// cholesky(Symmetric(::BandedMatrix)) (banded_chol!, src/symbanded/BandedCholesky.jl:2-13) and its ldiv! (:72-80).
// SURVEY.md 8(f) rank 3, the factorisation half.
//
// Storage (LAPACK symmetric band): 'U': A[i,k] (i <= k) at ab[(kd + i - k) + k*ldab]; 'L': A[i,k] (i >= k) at ab[(i - k) + k*ldab].
// Both are ONE upper factor U (A = U^T U; for 'L' the stored factor is L = U^T) addressed with two strides:
//     U(i,k) = p[i*si + k*sk]      'U': p = ab + kd, si = 1, sk = ldab-1        'L': p = ab, si = ldab-1, sk = 1
// so every kernel below is written once, on U(i,k), i <= k <= i + kd.
//
// kd <= 64 -- DPBTF2, the unblocked algorithm DPBTRF runs there (ILAENV gives NB = 1 for kd <= 64): per column j
//     d = sqrt(A[j,j]); row j *= 1/d (DSCAL by the reciprocal); trailing triangle -= x x^T, one FMA per entry with t = -x[c]
//     rounded first (DSYR as OpenBLAS runs it).  `pbtf2_window` keeps the kd+1 live columns in a shared-memory ring fed by
//     cp.async, every thread recomputes the reciprocal itself (no broadcast step), one barrier per column.  Each entry receives
//     its updates in ascending j exactly as the CPU does: factors are BIT-IDENTICAL to OpenBLAS dpbtrf_.
// kd > 64 -- blocked right-looking (the structure of DPBTRF, NB = 64 here): per panel of NB columns
//     `pb_potf2_reg`: the NB x NB diagonal block factored in registers by EVERY CTA of the grid, each of which then substitutes
//     U12 = U11^{-T} A12 for its 16 columns of A12 (CTA 0 instead writes the factored block back); `pb_syrk`: A22 -= U12^T U12
//     on the kd x kd window (64 x 64 tiles on DMMA.8x8x4).  The panel is (a device base counter) + (the launch's place in the
//     graph), so a CUDA graph of PB_GRAPH_PANELS panels is captured once per call and replayed; panels past the end are no-ops.  DPBTRF's DGEMM/DSYRK order is unspecified: compared to
//     OpenBLAS at 1e-13 * cond-ish tolerances and through ||U^T U - A||.
// pbtrs -- DPBTRS = two DTBSV sweeps per right-hand side ('U': U^T then U; 'L': L then L^T), run for ALL right-hand sides at
//     once (cluster pipeline of gbtrs_cluster.cu for the 'N' sweep, one chain block per right-hand side for the 'T' sweep).
#include <type_traits>

#include "common.cuh"

#ifndef PB_NB
#define PB_NB 64   // panel width of the blocked path (128 was measured: 719 against 557 ns/column -- the 1024-thread diagonal block spills at 64 registers)
#endif
#define PB_LG (PB_NB / 4)                 // lanes per row group in the register-resident diagonal block
#define PB_K1T (PB_LG * PB_LG)           // its thread count
#define PB_SLAB ((PB_NB + 4) * 68)       // doubles per staged U12 slab (either layout)
#define PB_PFD 8
#define PB_GRAPH_PANELS 128
#define PB_TRANSPOSE_KD 64  // dpbtrs: from this band width on, transpose the factor and run both sweeps as column sweeps

int bmb_cluster_solve(bmb200_ctx *h, int mode, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb);  // gbtrs_cluster.cu
int bmb_tbsv_t_multi(bmb200_ctx *h, int up, int unit, i64 n, i64 k, i64 nrhs, const double *dA, i64 lda, double *dB, i64 ldb);      // tb.cu

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(double *smem_dst, const double *gsrc, bool valid)  // !valid: writes 0.0, reads nothing
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 1/sqrt(a) to about an ulp: hardware seed + two Newton steps (~85 cycles against ~270 for sqrt followed by a divide)
__device__ __forceinline__ double pb_rsqrt(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = y * fma(-(h * y), y, 1.5);
    y = y * fma(-(h * y), y, 1.5);
    return y;
}

// d_state: [0] = info (0 ok, > 0 first non-positive pivot, 1-based); [1] is the panel counter of the blocked path.
// The kd+1 live columns sit in a shared-memory ring: entry U(k-d, k) at win[((k*P + d) & MASK)], P and the slot count powers of
// two, so one add and one mask address anything relative to the current column.  A thread owns up to E entries (r, c),
// 1 <= r <= c <= kd, of the trailing triangle RELATIVE to the current column j (offsets precomputed once, enumerated by
// ascending c so that the live entries of a shrinking triangle are a prefix).  One step: every thread loads the pivot, its
// entries and their two row-j factors, takes the square root / reciprocal itself (no broadcast), updates, one barrier.  The
// single warp per scheduler hides nothing, so the step is kept to the instructions the dependency chain needs.
template <int NT, int E>
__global__ void __launch_bounds__(NT)
pbtf2_window(i64 n, int kd, i64 si, i64 sk, double *__restrict__ p, int ring, int P, int *__restrict__ d_state, long long *__restrict__ stats)
{
    extern __shared__ double win[];
    const int MASK = ring * P - 1, tid = threadIdx.x;
    // entry offsets relative to column j: entry (r,c) -> c*P + (c-r); its factors S(j,j+r) -> r*P + r, S(j,j+c) -> c*P + c
    int oe[E], oxr[E], oxc[E];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const int e = tid + m * NT;
        int c = 1;
        while (c * (c + 1) / 2 <= e) ++c;
        const int r = e - c * (c - 1) / 2 + 1;
        oe[m] = c * P + (c - r);
        oxr[m] = r * P + r;
        oxc[m] = c * P + c;
    }
    const i64 diag = si + sk;
    const i64 tsk = (i64)tid * sk, tsi = (i64)tid * si;
    // column k of the band -> its ring slot; NT >= kd+1, so one copy per thread.  The running column advances by one per step:
    // pointer and ring offset are stepped, not recomputed (this runs inside the per-column chain of a single small CTA)
    i64 fk = 0;
    const double *fsrc = p - tsi;      // &U(k - tid, k) for k = 0
    int fbase = tid;                   // (k*P + tid) & MASK
    auto fetch_next = [&]() {
        if (fk < n && tid <= (fk < kd ? (int)fk : kd)) cp_async8(win + fbase, fsrc);
        cp_async_commit();
        ++fk;
        fsrc += diag;
        fbase = (fbase + P) & MASK;
    };
    for (int k = 0; k < kd + PB_PFD; ++k) fetch_next();
    cp_async_wait<0>();
    __syncthreads();
    int jP = 0;
    double *rowp = p + tsk;  // &U(j, j + tid)
    // one column step; FULL: the whole kd x kd trailing triangle is inside the matrix (all but the last kd columns), so which
    // entries a thread owns is a per-thread constant and nothing about the step depends on j
    bool mine[E];
#pragma unroll
    for (int m = 0; m < E; ++m) mine[m] = tid + m * NT < kd * (kd + 1) / 2;
    auto step = [&](auto full_tag, i64 j) -> bool {
        constexpr bool FULL = decltype(full_tag)::value;
        const int kn = FULL ? kd : (int)imin64_d(kd, n - 1 - j);
        const int cnt = kn * (kn + 1) / 2;
        // everything that does not depend on the pivot is loaded first
        const double ajj = win[jP];
        double ev[E], xr[E], xc[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const bool ok = FULL ? mine[m] : (tid + m * NT < cnt);
            ev[m] = ok ? win[(oe[m] + jP) & MASK] : 0.0;
            xr[m] = ok ? win[(oxr[m] + jP) & MASK] : 0.0;
            xc[m] = ok ? win[(oxc[m] + jP) & MASK] : 0.0;
        }
        const double rowv = (tid <= kn) ? win[(tid * P + tid + jP) & MASK] : 0.0;
        if (ajj <= 0.0) {  // not positive definite: DPBTF2 stops here with info = j+1 and the trailing window as updated so far
            for (int c = 0; c <= kn; ++c)
                for (int d = tid; d <= c; d += NT) p[(j + c - d) * si + (j + c) * sk] = win[(int)(((j + c) * P + d) & MASK)];
            if (tid == 0) d_state[0] = (int)(j + 1);
            cp_async_wait<0>();
            return false;
        }
        const double dj = sqrt(ajj), rinv = 1.0 / dj;  // DPBTF2: AJJ = SQRT(AJJ), then DSCAL by ONE / AJJ
#pragma unroll
        for (int m = 0; m < E; ++m) {
            if (FULL ? mine[m] : (tid + m * NT < cnt)) {
                const double t = -__dmul_rn(xc[m], rinv);
                if (t != 0.0) win[(oe[m] + jP) & MASK] = fma(t, __dmul_rn(xr[m], rinv), ev[m]);  // OpenBLAS dsyr skips zero entries of x
            }
        }
        if (tid <= kn) *rowp = tid == 0 ? dj : __dmul_rn(rowv, rinv);  // row j is final: U(j, j+tid)
        fetch_next();
        cp_async_wait<PB_PFD - 1>();
        __syncthreads();
        jP = (jP + P) & MASK;
        rowp += diag;
        return true;
    };
    i64 j = 0;
    for (; j + kd < n; ++j)
        if (!step(std::true_type{}, j)) return;
    for (; j < n; ++j)
        if (!step(std::false_type{}, j)) return;
    cp_async_wait<0>();
    (void)stats;
}

// ---- kd <= 31: DPBTF2 on ONE warp, register resident, no shared memory and no barrier.  Lane d owns diagonal d: it holds
// S(i, i+d) for the kd+1 live rows i = j .. j+kd in W[i mod (KD+1)] (the j loop is unrolled KD+1 times, so every register index is
// static) and, in Q, the same slots one window ahead (rows are fetched KD+1 steps before they are needed).  Step j: the pivot is
// lane 0's entry of row j (one shuffle), every lane takes sqrt and reciprocal itself, x_d = S(j,j+d)/sqrt is ALREADY in lane d;
// entry S(j+r, j+r+d) -= x_{r+d} x_r needs one broadcast (x_r) and one shift (x_{r+d}) per r.  Per column the dependency chain is
// shuffle + sqrt + divide + multiply + one shared-memory post + FMA; in practice the single warp is bound by its instruction
// count (~700 cycles per column at kd = 4, ~40 more per extra diagonal) and the trimmed window kernel is faster at every kd, so
// dispatch does not use it (tuning key pb_nodiag = -1 selects it; the tests do).  The
// order per entry is DPBTF2's (ascending j, t = -x_c rounded, zero x_c skipped): factors stay bit-identical to OpenBLAS.  After
// a non-positive pivot the warp goes `dead`: x = 0 (no update does anything), rows are written back as they stand.
template <int KD>
__global__ void __launch_bounds__(32)
pbtf2_diag(i64 n, int kd, i64 si, i64 sk, double *__restrict__ p0, int *__restrict__ d_state)
{
    constexpr int NS = KD + 1;
    const int lane = threadIdx.x;
    const i64 dstep = si + sk;                       // S(i,i+d) = p0[i*dstep + d*sk]
    double *base = p0 + (i64)lane * sk;
    const bool mine = lane <= kd;
    auto fetch = [&](i64 i) { return (mine && i + lane < n) ? base[i * dstep] : 0.0; };
    __shared__ double xs[2][64];
    xs[0][lane] = xs[1][lane] = 0.0;
    xs[0][32 + lane] = xs[1][32 + lane] = 0.0;
    __syncwarp();
    double W[NS], Q[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { W[s] = fetch(s); Q[s] = fetch(s + NS); }
    bool dead = false;
    int info = 0;
    for (i64 jb = 0; jb < n; jb += NS) {
#pragma unroll
        for (int ph = 0; ph < NS; ++ph) {
            const i64 j = jb + ph;
            if (j < n) {  // warp-uniform
                const double v = W[ph];
                const double ajj = __shfl_sync(0xffffffffu, v, 0);
                if (!dead && ajj <= 0.0) { dead = true; info = (int)(j + 1); }
                const double dj = sqrt(ajj), rinv = 1.0 / dj;
                const double x = dead ? 0.0 : __dmul_rn(v, rinv);
                const int kn = (int)imin64_d(kd, n - 1 - j);
                if (lane <= kn) base[j * dstep] = dead ? v : (lane == 0 ? dj : x);
                // every lane needs the whole scaled row: x_r (same for all lanes) and x_{r+d} (a shift).  One shared-memory post per
                // step (double-buffered by step parity, entries 32..63 stay 0) and two loads per r -- four SHFL per r from this single
                // warp measured ~70 cycles per r, batched or not (shuffles from one warp do not pipeline)
                double *xb = xs[ph & 1];
                xb[lane] = x;
                __syncwarp();
#pragma unroll
                for (int r = 1; r <= KD; ++r) {
                    const double xr = xb[r], xc = xb[lane + r];
                    const int slot = (ph + r) % NS;
                    if (lane + r <= kn && xc != 0.0) W[slot] = fma(-xc, xr, W[slot]);
                }
                W[ph] = Q[ph];
                Q[ph] = fetch(j + 2 * NS);
            }
        }
    }
    if (lane == 0 && info) d_state[0] = info;
}

// K1 of the blocked path: Cholesky of the NB x NB diagonal block, REGISTER resident.  Thread (a, b) = (tid / 16, tid % 16)
// owns the 4 x 4 patch rows 4a.., columns 4b.. (patches below the diagonal idle).  Step j = 4*jb + u (u unrolled, so every
// register index is static): the 16 lanes with a == jb sit in one warp -- the pivot comes by one shuffle from the diagonal
// lane, every lane takes the reciprocal square root itself, scales its part of row j (the other half-warp multiplies by 1) and
// posts it in shared memory; one barrier; every patch right/below subtracts x_r x_c.  Per step the dependency chain is
// shuffle + rsqrt + multiply + one shared-memory round trip + FMA (tools/fp64_issue.cu: ~30 + 72 + 10 + ~210 cycles); shared
// memory is addressed through precomputed 32-bit addresses and nothing inside the step touches global or constant memory.
__device__ __forceinline__ void pb_lds2(unsigned addr, double &x, double &y) { asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr)); }
__device__ __forceinline__ void pb_sts2(unsigned addr, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(x), "d"(y) : "memory"); }

#ifdef PB_DEBUG_TIMING
__device__ long long pb_stamps[4][8][2];
__device__ unsigned long long pb_lastend[4][2];  // latest end over ALL CTAs of the panel kernel / the update kernel
#define PB_STAMP_END(which, panel)                                                                  \
    do {                                                                                            \
        if (threadIdx.x == 0 && (panel) >= 10 && (panel) <= 13) {                                   \
            unsigned long long t_;                                                                  \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                  \
            atomicMax(&pb_lastend[(panel) - 10][which], t_);                                        \
        }                                                                                           \
    } while (0)
#define PB_STAMP(ev, panel)                                                                         \
    do {                                                                                            \
        if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && (panel) >= 10 && (panel) <= 13) {  \
            long long t_;                                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                  \
            pb_stamps[(panel) - 10][ev][blockIdx.x == 0 ? 0 : 1] = t_;                              \
        }                                                                                           \
        if (threadIdx.x == 0 && blockIdx.x == 0 && (panel) == 14 && ev == 4)                        \
            for (int p_ = 0; p_ < 4; ++p_)                                                          \
                for (int e_ = 0; e_ < 7; ++e_)                                                      \
                    printf("STAMP %d %d %lld %lld  last k1 %lld k3 %lld\n", p_ + 10, e_, pb_stamps[p_][e_][0] - pb_stamps[0][0][0], pb_stamps[p_][e_][1] - pb_stamps[0][0][0], (long long)pb_lastend[p_][0] - pb_stamps[0][0][0], (long long)pb_lastend[p_][1] - pb_stamps[0][0][0]); \
    } while (0)
#else
#define PB_STAMP(what, panel) do { } while (0)
#define PB_STAMP_END(which, panel) do { } while (0)
#endif
#define PB_TPC 16  // lanes per column of A12 in the substitution
// Since the fusion with the former substitution kernel this kernel is launched with one CTA per 16 columns of A12: EVERY CTA factors the diagonal block
// (redundantly -- the chain is the critical path anyway and the SMs are otherwise idle) and substitutes for its own columns row
// group by row group INSIDE the factorisation loop, from the rows the factorisation posts in shared memory: one launch, one
// dependent global round trip and the staging of U11 less per panel.  CTA 0 only factors and publishes: the factored block goes
// back into the matrix once every CTA of the grid has the unfactored block in its registers (arrival counter d_state[3]; only
// CTA 0 ever waits, so CTAs of a grid larger than one wave still get their turn).
__global__ void __launch_bounds__(PB_K1T)
pb_potf2_reg(i64 n_total, int kd, i64 si, i64 sk, double *__restrict__ p0, int *__restrict__ d_state, int q)
{
    __shared__ __align__(16) double xs[4][4][PB_NB];  // the four scaled rows of a row group, [row group mod 4][row][column]: a ring of four,
                                                      // because the substitution of a warp lags up to two row groups behind (below)
    __shared__ __align__(16) double rd[PB_NB + 2];
    __shared__ __align__(16) double dpatch[16];        // the diagonal 4 x 4 patch of the current row group
    __shared__ int s_fail;
    // programmatic dependent launch (the three kernels of a panel are captured with it): this grid was launched while its
    // predecessor was still running; nothing is read before the predecessor has completed and flushed, and the successor may
    // be launched right away so that its launch latency hides behind this grid
    // The panel is d_state[1] + q: q is this launch's place in the captured graph, d_state[1] the first panel of the replay --
    // written only by the LAST pb_syrk of the previous replay, so it may be read before the wait.  The info word is loaded
    // after the wait together with the block; the exit on it comes once those loads are in flight.
    int base;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(base) : "l"(d_state + 1));
    const int s_panel = base + q;
    const i64 j0 = (i64)s_panel * PB_NB;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (j0 >= n_total) return;
    int info0;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(info0) : "l"(d_state) : "memory");
    if (threadIdx.x == 0) s_fail = 0;
    PB_STAMP(0, s_panel);
    const int nbl = (int)imin64_d(PB_NB, n_total - j0);
    __syncthreads();
    double *p = p0 + j0 * (si + sk);  // U(j0, j0)
    // substitution phase, prepared now: CTA 0 only factors and publishes, CTA c >= 1 takes 16 columns of A12, 16 lanes per column,
    // lane tq holding the panel rows 4tq .. 4tq+3 of its column (they do not depend on the factorisation)
    const i64 c1 = j0 + nbl;                            // first column right of the panel
    const i64 ncols12 = imin64_d(kd, n_total - c1);     // columns of A12 (row j0+nbl-1 reaches c1-1+kd)
    const int tq = threadIdx.x % PB_TPC;
    const i64 tcol = ((i64)blockIdx.x - 1) * (PB_K1T / PB_TPC) + threadIdx.x / PB_TPC;
    const bool tlive = blockIdx.x > 0 && tcol < ncols12;
    const i64 tk = c1 + (tlive ? tcol : 0);
    const int ti0 = (int)imax64_d(0, tk - kd - j0);     // first stored panel row of column tk
    double ta[4];
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) {
        const int i = 4 * tq + s2;
        ta[s2] = (tlive && i >= ti0 && i < nbl) ? p0[(j0 + i) * si + tk * sk] : 0.0;
    }
    // the compiler otherwise re-reads %tid and rebuilds the shared-window base (S2R / S2UR, tens of cycles each) inside every
    // step, right on the dependency chain: read them once through volatile asm so that they have to stay in registers
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int a = tid / PB_LG, b = tid % PB_LG, lane = tid & 31, warp = tid >> 5;
    const bool upper = b >= a;
    double v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int r = 4 * a + u, c = 4 * b + w;
            v[u][w] = (upper && r <= c && c < nbl && c - r <= kd) ? p[(i64)r * si + (i64)c * sk] : ((r == c) ? 1.0 : 0.0);  // outside the band: 0, stays 0
        }
    unsigned xs0 = (unsigned)__cvta_generic_to_shared(&xs[0][0][0]);
    unsigned rd0 = (unsigned)__cvta_generic_to_shared(&rd[0]);
    unsigned fl0 = (unsigned)__cvta_generic_to_shared(&s_fail);
    unsigned dp0 = (unsigned)__cvta_generic_to_shared(&dpatch[0]);
    asm volatile("mov.u32 %0, %0;" : "+r"(dp0));
    asm volatile("mov.u32 %0, %0;" : "+r"(xs0));
    asm volatile("mov.u32 %0, %0;" : "+r"(rd0));
    asm volatile("mov.u32 %0, %0;" : "+r"(fl0));
    unsigned xr_a = xs0 + 32u * a, xc_a = xs0 + 32u * b;  // this patch's row / column factors inside one x buffer
    asm volatile("mov.u32 %0, %0;" : "+r"(xr_a));
    asm volatile("mov.u32 %0, %0;" : "+r"(xc_a));
    int failed = 0;
    // ---- substitution U12 = U11^{-T} A12 for this CTA's columns, one row group g (four rows) per call, INSIDE the factorisation
    // loop: the rows of U11 it needs are the four scaled rows just posted in xs (with rd, the reciprocal diagonal), and every
    // warp but the one on the critical path has slack until the next barrier.  Every lane runs the 4 x 4 forward substitution of
    // the group on its own registers (only the owner's, lane g of the column's 16, is meaningful), the owner's four results go
    // round by shuffles, the lanes below apply them to their four rows.  Per entry the subtractions come in ascending row order,
    // one FMA each -- the order of the column-at-a-time substitution.  The warp that holds the next row group skips its turn
    // (two iterations in a row) and catches up afterwards: hence the ring of four buffers.
    const bool subst_cta = blockIdx.x > 0 && ncols12 > 0;
    const unsigned tbase = (unsigned)lane & ~(unsigned)(PB_TPC - 1);
    int sg = 0;  // next row group this warp substitutes (warp-uniform)
    bool arrived = false;
    auto subst = [&](int g) {
        const unsigned gb = xs0 + (unsigned)(g & 3) * (4u * PB_NB * 8u);
        const unsigned dg = gb + (unsigned)g * 32u;  // row 0 of the buffer, column 4g
        double ul[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            pb_lds2(gb + r * (PB_NB * 8u) + 32u * tq, ul[r][0], ul[r][1]);
            pb_lds2(gb + r * (PB_NB * 8u) + 32u * tq + 16u, ul[r][2], ul[r][3]);
        }
        double d00, d01, d02, d03, d12, d13, d22, d23, r0, r1, r2, r3;
        pb_lds2(dg, d00, d01);
        pb_lds2(dg + 16u, d02, d03);
        pb_lds2(dg + PB_NB * 8u + 16u, d12, d13);
        pb_lds2(dg + 2u * PB_NB * 8u + 16u, d22, d23);
        pb_lds2(rd0 + 32u * g, r0, r1);
        pb_lds2(rd0 + 32u * g + 16u, r2, r3);
        (void)d00; (void)d22;
        const double x0 = __dmul_rn(ta[0], r0);
        const double x1 = __dmul_rn(fma(-d01, x0, ta[1]), r1);
        const double x2 = __dmul_rn(fma(-d12, x1, fma(-d02, x0, ta[2])), r2);
        const double x3 = __dmul_rn(fma(-d23, x2, fma(-d13, x1, fma(-d03, x0, ta[3]))), r3);
        const unsigned src = tbase + g;  // (posting the four results in shared memory instead of shuffling them measured the same)
        const double b0 = __shfl_sync(0xffffffffu, x0, src), b1 = __shfl_sync(0xffffffffu, x1, src);
        const double b2 = __shfl_sync(0xffffffffu, x2, src), b3 = __shfl_sync(0xffffffffu, x3, src);
        if (tq == g) { ta[0] = x0; ta[1] = x1; ta[2] = x2; ta[3] = x3; }
        if (tq > g) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double br = (r == 0) ? b0 : (r == 1) ? b1 : (r == 2) ? b2 : b3;
#pragma unroll
                for (int c = 0; c < 4; ++c) ta[c] = fma(-ul[r][c], br, ta[c]);
            }
        }
    };
    if (info0 != 0) return;  // an earlier panel was not positive definite
#pragma unroll 1
    for (int jb = 0; jb < PB_NB / 4 && 4 * jb < nbl; ++jb) {
        const bool inwarp = warp == ((jb * PB_LG) >> 5);  // this warp holds rows 4jb..4jb+3
        const bool rowgrp = a == jb && upper;           // this thread holds a piece of them
        const int dlane = (jb * PB_LG + jb) & 31;        // lane of the diagonal patch (a == b == jb)
        const bool bgt = b > jb;                         // this patch lies right of the diagonal patch
        // FOUR columns (one row group) per barrier.  The 4 x 4 diagonal patch D lives in ONE thread: its ten upper entries are
        // shuffled to every lane of the row-group warp (pipelined), every lane then factors D redundantly in registers -- four
        // rsqrt in sequence, no further communication -- and solves its own 4 x 4 patch of the row group against it (a 4 x 4
        // triangular substitution); the four scaled rows are posted and everyone below applies a rank-4 update.
        const unsigned buf = (unsigned)(jb & 3) * (4u * PB_NB * 8u);  // ring of four: later row groups post while this one is still read
        if (inwarp) {
            // the diagonal patch goes through shared memory (one post, broadcast loads): shuffles issued by a single warp do not
            // pipeline (ten 64-bit shuffles cost ~300 cycles here)
            double D[4][4];
            if (lane == dlane) {
#pragma unroll
                for (int r = 0; r < 4; ++r) { pb_sts2(dp0 + 32u * r, v[r][0], v[r][1]); pb_sts2(dp0 + 32u * r + 16u, v[r][2], v[r][3]); }
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 4; ++r) { pb_lds2(dp0 + 32u * r, D[r][0], D[r][1]); pb_lds2(dp0 + 32u * r + 16u, D[r][2], D[r][3]); }
            __syncwarp();
            double ri[4], U[4][4];
            int bad = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!bad && 4 * jb + u < nbl && !(D[u][u] > 0.0)) bad = 4 * jb + u + 1;  // not positive definite (NaN included)
                ri[u] = pb_rsqrt(D[u][u]);
#pragma unroll
                for (int c = u + 1; c < 4; ++c) U[u][c] = __dmul_rn(D[u][c], ri[u]);
#pragma unroll
                for (int r = u + 1; r < 4; ++r)
#pragma unroll
                    for (int c = r; c < 4; ++c) D[r][c] = fma(-U[u][r], U[u][c], D[r][c]);
            }
            if (bad && !failed) failed = bad;
            if (rowgrp && !failed) {
                double x[4][4];
                if (bgt) {
                    // a patch right of the diagonal patch: row u -= sum_{t<u} U[t][u] * x_t, then * ri[u]; every entry belongs to its row
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            double t = v[u][w];
#pragma unroll
                            for (int q = 0; q < u; ++q) t = fma(-U[q][u], x[q][w], t);
                            x[u][w] = __dmul_rn(t, ri[u]);
                            v[u][w] = x[u][w];
                        }
                } else {
                    // the diagonal patch itself: its factor is the U just computed (diagonal = D * rsqrt(D) = sqrt(D)); the posted
                    // rows carry zeros at and left of the diagonal
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            x[u][w] = (w > u) ? U[u][w] : 0.0;
                            if (w > u) v[u][w] = U[u][w];
                            else if (w == u) v[u][w] = __dmul_rn(D[u][u], ri[u]);
                        }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    pb_sts2(xc_a + buf + u * (PB_NB * 8u), x[u][0], x[u][1]);
                    pb_sts2(xc_a + buf + u * (PB_NB * 8u) + 16u, x[u][2], x[u][3]);
                }
            }
            if (lane == dlane) {
#pragma unroll
                for (int u = 0; u < 4; ++u) asm volatile("st.shared.f64 [%0], %1;" ::"r"(rd0 + 8u * (4 * jb + u)), "d"(ri[u]) : "memory");
                if (failed) asm volatile("st.shared.u32 [%0], %1;" ::"r"(fl0), "r"(failed) : "memory");
            }
        }
        __syncthreads();
        if (jb == 1 && tid == 0) {  // every thread has used what it loaded from the block (iteration 0): CTA 0 may overwrite it
            atomicAdd(&d_state[3], 1);
            arrived = true;
        }
        {
            int f;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(f) : "r"(fl0));
            if (f) { failed = f; break; }
        }
        // The warp that holds the NEXT row group applies the update first and alone -- its FMAs do not queue behind the other
        // seven warps' on the FP64 pipe -- and goes straight into its pivot phase; the others start when it has issued its
        // FMAs (named barrier 1: it arrives, they wait) and work in the shadow of its rsqrt chain.
        const bool has_next = jb + 1 < PB_NB / 4 && 4 * (jb + 1) < nbl;   // block-uniform
        const bool nextwarp = has_next && warp == (((jb + 1) * PB_LG) >> 5);
        if (has_next && !nextwarp) asm volatile("bar.sync 1, %0;" ::"n"(PB_K1T) : "memory");
        if (upper && a > jb) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                double xr[4], xc[4];
                pb_lds2(xr_a + buf + u * (PB_NB * 8u), xr[0], xr[1]);
                pb_lds2(xr_a + buf + u * (PB_NB * 8u) + 16u, xr[2], xr[3]);
                pb_lds2(xc_a + buf + u * (PB_NB * 8u), xc[0], xc[1]);
                pb_lds2(xc_a + buf + u * (PB_NB * 8u) + 16u, xc[2], xc[3]);
#pragma unroll
                for (int uu = 0; uu < 4; ++uu)
#pragma unroll
                    for (int w = 0; w < 4; ++w) v[uu][w] = fma(-xr[uu], xc[w], v[uu][w]);
            }
        }
        if (nextwarp) asm volatile("bar.arrive 1, %0;" ::"n"(PB_K1T) : "memory");
        if (subst_cta && !nextwarp)  // at most two groups per turn: a warp that owes three spreads them over two turns
            for (int turn = 0; turn < 2 && sg <= jb; ++turn) subst(sg++);
    }
    __syncthreads();
    PB_STAMP(1, s_panel);
    if (tid == 0 && !arrived) atomicAdd(&d_state[3], 1);  // a one-iteration panel, or an exit from iteration 0
    if (!failed) {
        int f;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(f) : "r"(fl0));
        failed = f;
    }
    auto publish = [&]() {  // CTA 0: the block as it stands in the registers goes back into the matrix
        if (tid == 0) {
            const volatile int *cnt = d_state + 3;
            while (*cnt < (int)gridDim.x) { }
            __threadfence();
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int r = 4 * a + u, c = 4 * b + w;
                if (upper && r <= c && c < nbl && c - r <= kd) p[(i64)r * si + (i64)c * sk] = v[u][w];
            }
    };
    if (failed) {  // not positive definite: the block as updated so far goes back (CTA 0), the panel counter stops
        if (blockIdx.x == 0) {
            if (tid == 0) d_state[0] = (int)(j0 + failed);
            publish();
        }
        return;
    }
    if (blockIdx.x == 0) {  // the publisher: its write-back overlaps the other CTAs' substitution
        publish();
        PB_STAMP(3, s_panel);
        return;
    }
    if (ncols12 <= 0) return;
    {
        // the row groups this warp still owes (it was on the critical path at the end), then the result
        const int ngroups = (nbl + 3) / 4;
        while (sg < ngroups) subst(sg++);
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
            const int i = 4 * tq + s2;
            if (tlive && i >= ti0 && i < nbl) p0[(j0 + i) * si + tk * sk] = ta[s2];
        }
    }
    PB_STAMP(2, s_panel);
    PB_STAMP_END(0, s_panel);
}

// K3: A22(r,c) -= sum_i U12(i,r) U12(i,c) for r <= c inside the window of kd columns right of the panel.  One CTA per 64 x 64
// tile of the upper triangle; the two 64-column slabs of U12 are staged in shared memory (zero outside the band / the matrix)
// and the product runs on the FP64 tensor cores (DMMA.8x8x4): with one CTA per SM there are only two warps per scheduler,
// which saturates the tensor pipe but leaves the FP64 FMA pipe at a fraction of its rate (measured: 16 FMAs per thread took
// ~700 cycles in the first version of this kernel).  Warp w owns the eight 8 x 8 tiles of tile row w: one A fragment and eight
// B fragments per k-step of 4.  blockIdx.x enumerates the tiles (tr <= tc) of the largest window; tiles past the actual window
// return.  Staged element (panel row i, window column x) sits at s[i*SI + x*SX]: the dimension that is contiguous in global
// memory is contiguous in shared memory, and the other pitch is chosen so that a fragment load (4 values of i x 8 values of x)
// is conflict-free per half-warp (16 doubles in 16 different 8-byte banks): SX = NB + 4 ('U'), SI = 68 ('L'), both 4 mod 16.
__device__ __forceinline__ void pb_dmma884(double &d0, double &d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256)
pb_syrk(i64 n, int kd, i64 si, i64 sk, double *__restrict__ p0, int *__restrict__ d_state, int ntile1d, int q, int bump, int bulk_ok, int c_early)
{
    extern __shared__ __align__(16) double pb_sm[];
    __shared__ __align__(8) unsigned long long tma_bar;  // completion of the bulk copies of this CTA's slabs
    double *sr = pb_sm, *sc = pb_sm + PB_SLAB;
    int base;  // see pb_potf2_reg
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(base) : "l"(d_state + 1));
    const int panel = base + q;
    const i64 j0 = (i64)panel * PB_NB;
    // everything that does not touch the predecessor's data comes before the wait: the tile, which of its slabs are whole (all
    // NB x 64 entries inside the band and the matrix -- those arrive as 64 bulk copies of 512 contiguous bytes each, one per
    // slab column ('U') or row ('L'), completing on an mbarrier), and the barrier's transaction count
    const int tid = threadIdx.x;
    const int nbl = (int)imin64_d(PB_NB, n - j0);
    const i64 c1 = j0 + nbl;
    const i64 ncols = imin64_d(kd, n - c1);
    int tc = 0, rem = blockIdx.x;  // tile index -> (tr, tc), tr <= tc < ntile1d
    while (rem > tc) { rem -= tc + 1; ++tc; }
    const int tr = rem;
    auto whole = [&](int tile) { return bulk_ok && nbl == PB_NB && (i64)tile * 64 + 64 <= ncols && tile * 64 + 127 <= kd; };
    const bool bulk_r = whole(tr), bulk_c = tc != tr && whole(tc);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&tma_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const unsigned bytes = ((bulk_r ? 1u : 0u) + (bulk_c ? 1u : 0u)) * (unsigned)(PB_NB * 64 * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    }
    __syncthreads();
    // the C tile too: the window was last written by the update kernel of the PREVIOUS panel, which had completed before the
    // panel kernel this grid waits for could pass its own wait -- the panel kernel itself writes rows j0 .. j0+NB-1 only
    const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const i64 rr = (i64)tr * 64 + warp * 8 + lr;  // this lane's window row (accumulator layout of DMMA.8x8x4, below)
    double cv[8][2];
    auto load_c = [&]() {
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
                const i64 cc = (i64)tc * 64 + t * 8 + 2 * lc + q2;
                cv[t][q2] = (j0 < n && rr <= cc && cc < ncols) ? p0[(c1 + rr) * si + (c1 + cc) * sk] : 0.0;
            }
    };
    if (c_early) load_c();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (j0 >= n) return;
    int info0;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(info0) : "l"(d_state) : "memory");
    PB_STAMP(4, panel);
    if (!c_early) load_c();
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the arrival counter of the panel kernel back to zero; the last panel of a replay
        d_state[3] = 0;                         // moves the base on for the next one
        if (bump) d_state[1] = base + bump;
    }
    if (ncols <= 0) return;
    if ((i64)tc * 64 >= ncols) return;
    double *p = p0;
    const int SI = (si == 1) ? 1 : 68, SX = (si == 1) ? PB_NB + 4 : 1;
    auto stage = [&](double *s, int tile) {
        for (int e = tid; e < PB_NB * 64; e += 256) {
            int i, x;
            if (si == 1) { i = e % PB_NB; x = e / PB_NB; } else { x = e % 64; i = e / 64; }
            const i64 cc = (i64)tile * 64 + x, k = c1 + cc;
            const bool ok = cc < ncols && i < nbl && k <= j0 + i + kd;
            cp_async8_zfill(s + i * SI + x * SX, p + (j0 + (ok ? i : 0)) * si + (ok ? k : j0) * sk, ok);
        }
    };
    auto stage_bulk = [&](double *s, int tile, int c) {  // copy c of 64: slab column c ('U', si == 1) or slab row c ('L', sk == 1)
        const double *src = p + j0 * si + (c1 + (i64)tile * 64) * sk + (i64)c * ((si == 1) ? sk : si);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s + c * (PB_NB + 4));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                     "r"((unsigned)(64 * sizeof(double))), "r"(bar)
                     : "memory");
    };
    if (bulk_r) { if (tid < 64) stage_bulk(sr, tr, tid); } else stage(sr, tr);
    if (tc != tr) {
        if (bulk_c) { if (tid >= 64 && tid < 128) stage_bulk(sc, tc, tid - 64); } else stage(sc, tc);
    }
    cp_async_commit();
    // accumulator layout of DMMA.8x8x4: lane holds C[row = lane/4][col = 2*(lane%4) + {0,1}] of each 8 x 8 tile
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    cp_async_wait<0>();
    if (bulk_r || bulk_c) {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    __syncthreads();
    if (info0 != 0) return;  // the panel was not positive definite: nothing is written
    PB_STAMP(5, panel);
    const double *scc = (tc != tr) ? sc : sr;
    const double *ap = sr + lc * SI + (warp * 8 + lr) * SX;  // A fragment: A[row = lane/4][k = lane%4] = U12(k0 + k, row)
    const double *bp = scc + lc * SI + lr * SX;              // B fragment: B[k = lane%4][col = lane/4] = U12(k0 + k, col)
#pragma unroll 4
    for (int k0 = 0; k0 < PB_NB; k0 += 4) {
        const double av = ap[k0 * SI];
        double bv[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) bv[t] = bp[k0 * SI + t * 8 * SX];
#pragma unroll
        for (int t = 0; t < 8; ++t) pb_dmma884(acc[t][0], acc[t][1], av, bv[t]);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const i64 cc = (i64)tc * 64 + t * 8 + 2 * lc + q;
            if (rr <= cc && cc < ncols) p[(c1 + rr) * si + (c1 + cc) * sk] = cv[t][q] - acc[t][q];
        }
    PB_STAMP_END(1, panel);
    PB_STAMP(6, panel);
}

// ---- helpers of the narrow-band dpbtrs: both sweeps run through the tuned multi-RHS back substitution of bmb200_dgbtrs ----
// A lower-triangular solve becomes an upper-triangular one under the reversal i -> n-1-i.  `mode` 0: dst = reversal of U^T
// ('U' factor, first sweep), 1: dst = reversal of L ('L' factor, first sweep); dst is 'U' triangular-band storage, ld = kd+1.
__global__ void __launch_bounds__(256)
pb_reverse_factor(int mode, i64 n, int kd, const double *__restrict__ src, i64 lds, double *__restrict__ dst)
{
    const i64 total = n * (kd + 1);
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 jp = e / (kd + 1);
        const int b = (int)(e - jp * (kd + 1));
        double v = 0.0;
        if (jp - (kd - b) >= 0) {  // in-matrix entry M[i', j'] with i' = j' - (kd - b)
            if (mode == 0) v = src[b + (n - 1 - jp + (kd - b)) * lds];   // U[n-1-j', n-1-i'] : band row b of column n-1-i'
            else v = src[(kd - b) + (n - 1 - jp) * lds];                 // L[n-1-i', n-1-j'] : band row kd-b of column n-1-j'
        }
        dst[e] = v;
    }
}
__global__ void pb_iota(i64 n, i64 *__restrict__ p)
{
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = i + 1;
}
__global__ void pb_reverse_rows(i64 n, i64 nrhs, double *__restrict__ b, i64 ldb)
{
    const i64 half = n / 2, total = half * nrhs;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 q = e / half, i = e - q * half;
        double *c = b + q * ldb;
        const double t = c[i];
        c[i] = c[n - 1 - i];
        c[n - 1 - i] = t;
    }
}

// Non-unit triangular band solve T x = b ('N') for narrow bands through the multi-RHS back substitution of bmb200_dgbtrs (kl = 0,
// identity pivots): 'U' directly -- DGBTRS' U sweep IS dtbsv('U','N','N') -- and 'L' as the upper-triangular solve of the
// reversed system (same operations in the same order per entry: bit-identical to dtbsv('L','N','N')).  Used by bmb200_dtbsv.
__global__ void pb_unit_diag(i64 n, int k, double *__restrict__ m)  // 'U' storage, ld = k+1: the diagonal is band row k
{
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) m[k + j * (k + 1)] = 1.0;
}
__global__ void pb_copy_band(i64 n, int rows, const double *__restrict__ src, i64 lds, double *__restrict__ dst)
{
    const i64 total = n * rows;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 j = e / rows;
        dst[e] = src[(e - j * rows) + j * lds];
    }
}

int bmb_tri_solve_via_gbtrs(bmb200_ctx *h, int up, int tr, int unit, i64 n, i64 k, i64 nrhs, const double *dA, i64 lda, double *dB, i64 ldb)
{
    // op(T) upper: ('U','N') as stored; ('L','T') after a band transpose.  op(T) lower: reversed system -- ('L','N') from L,
    // ('U','T') from U^T (gather modes 1 / 0 of pb_reverse_factor).  The transposed forms run as column sweeps: equal to the
    // dot-product dtbsv('T') to rounding (1e-13), the 'N' forms bit for bit.
    const bool direct = up && !tr && !unit;  // a unit diagonal is made explicit in a copy (dividing by 1.0 is exact)
    const size_t fbytes = direct ? 0 : (size_t)n * (size_t)(k + 1) * sizeof(double), need = fbytes + (size_t)n * sizeof(i64);
    if (need > h->backup_bytes) {
        if (h->backup) { cudaStreamSynchronize(h->stream); cudaFree(h->backup); h->backup = nullptr; h->backup_bytes = 0; }
        BMB_CUDA(h, cudaMalloc(&h->backup, need));
        h->backup_bytes = need;
    }
    double *M = (double *)h->backup;
    i64 *idp = (i64 *)((char *)h->backup + fbytes);
    pb_iota<<<(unsigned)imin64(cdiv64(n, 256), (i64)h->sm_count * 8), 256, 0, h->stream>>>(n, idp);
    BMB_LAUNCH_CHECK(h);
    if (direct) return bmb200_dgbtrs(h, 'N', n, 0, k, nrhs, dA, lda, idp, dB, ldb);
    const unsigned gb = (unsigned)imin64(cdiv64(n * (k + 1), 256), (i64)h->sm_count * 16);
    const unsigned gd = (unsigned)imin64(cdiv64(n, 256), (i64)h->sm_count * 8);
    if ((up != 0) != (tr != 0)) {  // op(T) is upper triangular: ('U','N') with a unit diagonal, or ('L','T') = L^T after the band transpose
        if (up) {
            pb_copy_band<<<gb, 256, 0, h->stream>>>(n, (int)k + 1, dA, lda, M);
            BMB_LAUNCH_CHECK(h);
        } else {
            const int rc = bmb200_dband_transpose(h, n, n, k, 0, dA, lda, M, k + 1);
            if (rc) return rc;
        }
        if (unit) { pb_unit_diag<<<gd, 256, 0, h->stream>>>(n, (int)k, M); BMB_LAUNCH_CHECK(h); }
        return bmb200_dgbtrs(h, 'N', n, 0, k, nrhs, M, k + 1, idp, dB, ldb);
    }
    const unsigned gr = (unsigned)imin64(cdiv64(imax64(1, (n / 2) * nrhs), 256), (i64)h->sm_count * 16);
    pb_reverse_factor<<<gb, 256, 0, h->stream>>>(up ? 0 : 1, n, (int)k, dA, lda, M);
    if (unit) pb_unit_diag<<<gd, 256, 0, h->stream>>>(n, (int)k, M);
    pb_reverse_rows<<<gr, 256, 0, h->stream>>>(n, nrhs, dB, ldb);
    h->launches += 2 + (unit ? 1 : 0);
    BMB_CUDA(h, cudaGetLastError());
    const int rc = bmb200_dgbtrs(h, 'N', n, 0, k, nrhs, M, k + 1, idp, dB, ldb);
    if (rc) return rc;
    pb_reverse_rows<<<gr, 256, 0, h->stream>>>(n, nrhs, dB, ldb);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// The transposed non-unit solve for WIDE bands: the factor is transposed once on the device (one HBM pass into grow-only
// workspace) and the sweep runs as a column sweep through the cluster pipeline (the chain of n dependent dot products of the
// dtbsv('T') kernel costs ~2 us per column at k = 1024).  Returns 1 if the cluster pipeline does not take the shape.
int bmb_tri_solve_transposed_wide(bmb200_ctx *h, int up, int unit, i64 n, i64 k, i64 nrhs, const double *dA, i64 lda, double *dB, i64 ldb)
{
    const size_t need = (size_t)n * (size_t)(k + 1) * sizeof(double);
    if (need > h->backup_bytes) {
        if (h->backup) { cudaStreamSynchronize(h->stream); cudaFree(h->backup); h->backup = nullptr; h->backup_bytes = 0; }
        BMB_CUDA(h, cudaMalloc(&h->backup, need));
        h->backup_bytes = need;
    }
    double *tr = (double *)h->backup;
    const int rc = bmb200_dband_transpose(h, n, n, up ? 0 : k, up ? k : 0, dA, lda, tr, k + 1);
    if (rc) return rc;
    // U^T = lower triangular ('L' storage: mode 3 dividing, 2 unit); L^T = upper triangular ('U' storage: mode 0 dividing, 1 unit)
    return up ? bmb_cluster_solve(h, unit ? 2 : 3, n, k, 0, nrhs, tr, k + 1, dB, ldb) : bmb_cluster_solve(h, unit ? 1 : 0, n, 0, k, nrhs, tr, k + 1, dB, ldb);
}

// ---- ldiv!(transpose(F), B) (src/banded/linalg.jl:41-47 -> dgbtrs_('T')): U^T y = b, then L^T x = y with the interchanges undone.
// The U^T sweep never involves the pivots: it always runs as a column sweep on a reversed / transposed copy of U (helpers
// above).  For interchange-free factors L^T is a unit upper-triangular band: gathered once into 'U' storage with an explicit unit
// diagonal and solved the same way (dividing by 1 is exact).  *u_done / *l_done tell the caller (gbtrs.cu) which halves are left
// for its generic kernel. ----
__global__ void __launch_bounds__(256)
pb_gather_lt(i64 n, int kl, int kv, const double *__restrict__ ab, i64 ldab, double *__restrict__ dst)
{
    const i64 total = n * (kl + 1);
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 j = e / (kl + 1);
        const int b = (int)(e - j * (kl + 1));  // 'U' storage of L^T: band row b holds (L^T)[j-(kl-b), j] = L[j, j-(kl-b)]
        const i64 i = j - (kl - b);
        double v = 0.0;
        if (b == kl) v = 1.0;
        else if (i >= 0) v = ab[(kv + (kl - b)) + i * ldab];
        dst[e] = v;
    }
}
int bmb_count_interchanges(bmb200_ctx *h, i64 n, const i64 *d_ipiv, int *count);  // gbtrs_blocked.cu

int bmb_gbtrs_t_fast(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv, double *dB, i64 ldb, int *u_done,
                     int *l_done)
{
    *u_done = *l_done = 0;
    if (n < 2) return 0;
    const i64 kv = kl + ku;
    int rc = (kv <= 63) ? bmb_tri_solve_via_gbtrs(h, 1, 1, 0, n, kv, nrhs, dAB, ldab, dB, ldb) : bmb_tri_solve_transposed_wide(h, 1, 0, n, kv, nrhs, dAB, ldab, dB, ldb);
    if (rc == 1) return 0;  // the cluster pipeline does not take this shape: everything is left to the caller
    if (rc) return rc;
    *u_done = 1;
    if (kl == 0) { *l_done = 1; return 0; }
    int hc = 0;
    rc = bmb_count_interchanges(h, n, d_ipiv, &hc);
    if (rc) return rc;
    if (hc != 0) return 0;  // interchanges: the L^T half stays with the caller's kernel
    const size_t fbytes = (size_t)n * (size_t)(kl + 1) * sizeof(double), need = fbytes + (size_t)n * sizeof(i64);
    if (need > h->backup_bytes) {
        if (h->backup) { cudaStreamSynchronize(h->stream); cudaFree(h->backup); h->backup = nullptr; h->backup_bytes = 0; }
        BMB_CUDA(h, cudaMalloc(&h->backup, need));
        h->backup_bytes = need;
    }
    double *M = (double *)h->backup;
    i64 *idp = (i64 *)((char *)h->backup + fbytes);
    pb_gather_lt<<<(unsigned)imin64(cdiv64(n * (kl + 1), 256), (i64)h->sm_count * 16), 256, 0, h->stream>>>(n, (int)kl, (int)kv, dAB, ldab, M);
    BMB_LAUNCH_CHECK(h);
    if (kl <= 63) {
        pb_iota<<<(unsigned)imin64(cdiv64(n, 256), (i64)h->sm_count * 8), 256, 0, h->stream>>>(n, idp);
        BMB_LAUNCH_CHECK(h);
        rc = bmb200_dgbtrs(h, 'N', n, 0, kl, nrhs, M, kl + 1, idp, dB, ldb);
    } else {
        rc = bmb_cluster_solve(h, 0, n, 0, kl, nrhs, M, kl + 1, dB, ldb);
        if (rc == 1) return 0;
    }
    if (rc) return rc;
    *l_done = 1;
    return 0;
}

static int pb_check(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, int64_t ldab, int &up)
{
    if (!h) return -1;
    up = (uplo == 'U' || uplo == 'u');
    if (!up && !(uplo == 'L' || uplo == 'l')) return -2;
    if (n < 0) return -3;
    if (kd < 0) return -4;
    return 0;
}

extern "C" int bmb200_dpbtrf(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, double *dAB, int64_t ldab, int *info)
{
    int up;
    const int rc0 = pb_check(h, uplo, n, kd, ldab, up);
    if (rc0) return rc0;
    if (ldab < kd + 1) return -6;
    if (!info) return -7;
    *info = 0;
    if (n == 0) return 0;
    if (!dAB) return -5;
    if (kd >= ((int64_t)1 << 24)) return -4;
    DeviceGuard g(h->device);
    const i64 si = up ? 1 : ldab - 1, sk = up ? ldab - 1 : 1;
    double *p0 = dAB + (up ? kd : 0);        // 'U': the diagonal lives in band row kd
    if (kd > n - 1) kd = n > 1 ? n - 1 : 0;  // bands beyond the matrix are never referenced (LAPACK: kn = min(kd, n-j))
    int *d_state = h->d_info + 24;
    const int init[4] = {0, 0, 0, 0};  // info, first panel of the graph replay, -, arrival counter of the panel kernel
    BMB_CUDA(h, cudaMemcpyAsync(d_state, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    const bool blocked = kd > 64;
    long long *dstats = nullptr;  // (the cycle breakdown that guided the window kernel was removed with its clock reads)
    // measured (n = 2^19, ns per column, one-warp register kernel / window kernel after its per-step instruction count was cut):
    // kd = 2: 318 / 254, 4: 343 / 254, 8: 463 / 276, 16: 610 / 358, 31: 1222 / 484 -- the window kernel takes every kd <= 64; the
    // one-warp kernel stays reachable through the tuning block (pb_nodiag = -1) and is tested on every kd <= 31
    if (!blocked && kd <= 31 && h->tune.pb_nodiag == -1) {
        if (kd <= 4) pbtf2_diag<4><<<1, 32, 0, h->stream>>>(n, (int)kd, si, sk, p0, d_state);
        else if (kd <= 8) pbtf2_diag<8><<<1, 32, 0, h->stream>>>(n, (int)kd, si, sk, p0, d_state);
        else if (kd <= 16) pbtf2_diag<16><<<1, 32, 0, h->stream>>>(n, (int)kd, si, sk, p0, d_state);
        else pbtf2_diag<31><<<1, 32, 0, h->stream>>>(n, (int)kd, si, sk, p0, d_state);
        BMB_LAUNCH_CHECK(h);
    } else if (!blocked) {
        const int kdw = (int)kd;
        int ring = 16, P = 8;
        while (ring < kdw + 1 + PB_PFD) ring <<= 1;
        while (P < kdw + 1) P <<= 1;  // powers of two: ring addressing is an add and a mask
        const size_t smem = (size_t)ring * P * sizeof(double);
        typedef void (*k1_t)(i64, int, i64, i64, double *, int, int, int *, long long *);
        // threads >= kd+1 (one row entry each) and threads * E >= kd(kd+1)/2 (the trailing triangle)
        k1_t k1;
        unsigned nt1;
        // the fewest entry slots E per thread for the thread count (every slot costs instructions on the per-column chain)
        if (kdw <= 7) { k1 = pbtf2_window<32, 1>; nt1 = 32; }          // T = kd(kd+1)/2 <= 28
        else if (kdw <= 10) { k1 = pbtf2_window<32, 2>; nt1 = 32; }    // <= 55
        else if (kdw <= 15) { k1 = pbtf2_window<64, 2>; nt1 = 64; }    // <= 120
        else if (kdw <= 22) { k1 = pbtf2_window<128, 2>; nt1 = 128; }  // <= 253
        else if (kdw <= 31) { k1 = pbtf2_window<128, 4>; nt1 = 128; }  // <= 496
        else if (kdw <= 44) { k1 = pbtf2_window<256, 4>; nt1 = 256; }  // <= 990
        else if (kdw <= 63) { k1 = pbtf2_window<256, 8>; nt1 = 256; }  // <= 2016
        else { k1 = pbtf2_window<256, 9>; nt1 = 256; }                 // 2080
        BMB_CUDA(h, cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1<<<1, nt1, smem, h->stream>>>(n, (int)kd, si, sk, p0, ring, P, d_state, dstats);
        BMB_LAUNCH_CHECK(h);
    } else {
        const i64 npanels = cdiv64(n, PB_NB);
        const int ntile1d = (int)cdiv64(imin64(kd, n), 64);
        // whole slabs of U12 reach shared memory as bulk copies (TMA) when their 512-byte pieces are 16-byte aligned: the pieces
        // start at even element offsets (j0, c1 and the tile origin are multiples of 64) times the other stride
        const int bulk_ok = (h->tune.pb_nobulk == 0 && ((uintptr_t)p0 % 16) == 0 && (((si == 1) ? sk : si) % 2) == 0) ? 1 : 0;
        const unsigned ntiles = (unsigned)(ntile1d * (ntile1d + 1) / 2);
        const unsigned trsm_blocks = (unsigned)cdiv64(imin64(kd, n) * PB_TPC, 256);
        const size_t smem3 = (size_t)2 * PB_SLAB * sizeof(double);
        BMB_CUDA(h, cudaFuncSetAttribute(pb_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        // one graph of PB_GRAPH_PANELS panels (panel kernel + update kernel each), replayed; panel = d_state[1] + place in the graph
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const i64 chunk = imin64(npanels, PB_GRAPH_PANELS);
        // the caller's stream may be the legacy default stream, which cannot be captured: capture and replay on the handle's
        // own stream, ordered after the caller's stream by an event (the call synchronises before it returns)
        cudaStream_t gs = h->copy_stream;
        BMB_CUDA(h, cudaEventRecord(h->ev[3], h->stream));
        BMB_CUDA(h, cudaStreamWaitEvent(gs, h->ev[3], 0));
        cudaError_t ce = cudaSuccess;
        for (int pdl = h->tune.pb_nopdl ? 0 : 1; pdl >= 0; --pdl) {
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            auto cfg = [&](unsigned grid, unsigned block, size_t sm) {
                cudaLaunchConfig_t c = {};
                c.gridDim = dim3(grid);
                c.blockDim = dim3(block);
                c.dynamicSmemBytes = sm;
                c.stream = gs;
                c.attrs = attr;
                c.numAttrs = pdl ? 1 : 0;
                return c;
            };
            BMB_CUDA(h, cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
            ce = cudaSuccess;
            for (i64 q = 0; q < chunk && ce == cudaSuccess; ++q) {
                cudaLaunchConfig_t c1c = cfg(trsm_blocks + 1, PB_K1T, 0), c3c = cfg(ntiles, 256, smem3);
                ce = cudaLaunchKernelEx(&c1c, pb_potf2_reg, (i64)n, (int)kd, si, sk, p0, d_state, (int)q);
                if (ce == cudaSuccess) ce = cudaLaunchKernelEx(&c3c, pb_syrk, (i64)n, (int)kd, si, sk, p0, d_state, ntile1d, (int)q, (int)(q == chunk - 1 ? chunk : 0), bulk_ok, (int)(h->tune.pb_clate ? 0 : 1));
            }
            const cudaError_t ee = cudaStreamEndCapture(gs, &graph);
            if (ce == cudaSuccess) ce = ee;
            if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
            if (ce == cudaSuccess) break;
            if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
            exec = nullptr;
            (void)cudaGetLastError();
        }
        if (ce != cudaSuccess) { snprintf(h->err, sizeof(h->err), "dpbtrf: graph capture / instantiation failed: %s", cudaGetErrorString(ce)); return BMB200_ERR_CUDA - (int)ce; }
        for (i64 q = 0; q < npanels; q += chunk) {
            ce = cudaGraphLaunch(exec, gs);
            if (ce != cudaSuccess) break;
        }
        h->launches += 2 * cdiv64(npanels, chunk) * chunk;
        const cudaError_t se = cudaStreamSynchronize(gs);
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        BMB_CUDA(h, ce);
        BMB_CUDA(h, se);
    }
    int st[2];
    BMB_CUDA(h, cudaMemcpyAsync(st, d_state, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    *info = st[0];
    return 0;
}

extern "C" int bmb200_dpbtrs(bmb200_handle_t h, char uplo, int64_t n, int64_t kd, int64_t nrhs, const double *dAB, int64_t ldab, double *dB,
                             int64_t ldb)
{
    int up;
    const int rc0 = pb_check(h, uplo, n, kd, ldab, up);
    if (rc0) return rc0;
    if (nrhs < 0) return -5;
    if (ldab < kd + 1) return -7;
    if (ldb < (n > 1 ? n : 1)) return -9;
    if (n == 0 || nrhs == 0) return 0;
    if (!dAB) return -6;
    if (!dB) return -8;
    DeviceGuard g(h->device);
    // DPBTRS: 'U': solve U^T y = b, then U x = y;  'L': solve L y = b, then L^T x = y  (DTBSV per right-hand side), every
    // sweep as a COLUMN sweep for all right-hand sides at once:
    //   kd >= 64: the cluster pipeline of gbtrs_cluster.cu; the transposed sweep runs on a transposed copy of the factor (one
    //             HBM pass; the chain of n dependent dot products of dtbsv('T') costs ~2 us per column at kd = 1024);
    //   kd <  64: the multi-RHS back substitution of bmb200_dgbtrs (kl = 0, identity pivots: ~45 ns per column for all
    //             right-hand sides together); lower-triangular sweeps are the upper-triangular ones of the reversed system.
    // Equal to DPBTRS to rounding (the 'T' sweeps of the reference are dot-product forms).
    int rc;
    if (kd >= PB_TRANSPOSE_KD) {
        if (up) {
            rc = bmb_tri_solve_transposed_wide(h, 1, 0, n, kd, nrhs, dAB, ldab, dB, ldb);
            if (rc == 0) rc = bmb_cluster_solve(h, 0, n, 0, kd, nrhs, dAB, ldab, dB, ldb);
        } else {
            rc = bmb_cluster_solve(h, 3, n, kd, 0, nrhs, dAB, ldab, dB, ldb);
            if (rc == 0) rc = bmb_tri_solve_transposed_wide(h, 0, 0, n, kd, nrhs, dAB, ldab, dB, ldb);
        }
    } else {
        rc = bmb_tri_solve_via_gbtrs(h, up, up ? 1 : 0, 0, n, kd, nrhs, dAB, ldab, dB, ldb);
        if (rc == 0) rc = bmb_tri_solve_via_gbtrs(h, up, up ? 0 : 1, 0, n, kd, nrhs, dAB, ldab, dB, ldb);
        return rc;
    }
    if (rc == 1) {
        snprintf(h->err, sizeof(h->err), "dpbtrs: band width %lld is not supported by the cluster pipeline on this device", (long long)kd);
        return BMB200_ERR_CUDA;
    }
    return rc;
}
