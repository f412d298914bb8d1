// gbtrs_cluster.cu -- wide-band solve with interchange-free factors (ipiv = 1:n), one thread-block CLUSTER per
// right-hand side, pipelined through distributed shared memory.
//
// Replaces the one-CTA-per-RHS sweep of gbtrs_blocked.cu for LAPACK.gbtrs! (src/banded/linalg.jl:28) when the pivot
// vector is the identity (every diagonally dominant system, e.g. examples/finitedifference_2d.jl = BASELINE config C5).
// One SM can pull only ~64 B/clk of factor entries out of L2, which bounded the single-CTA kernel (6 us per 16-column
// panel at l=u=1024).  Here the band is cut, per 16-row block, into a NEAR part (the GC_D row blocks next to the
// diagonal block) and a FAR part:
//   * the LEADER CTA (cluster rank 0) owns the dependency chain.  Warp w takes the 16-row blocks t = w, w+GC_LW, ...: it
//     receives the block from a worker (hand-off ring in its shared memory), applies the updates of the last GC_D
//     panels as their solutions appear, solves the 16 x 16 diagonal triangle with shuffles, publishes the 16 solution
//     entries in its own shared memory (next warp's input: plain doubles + an mbarrier arrive), in global memory (the result)
//     and in every worker's shared memory (DSMEM stores of self-validating 16-byte cells, no flags or fences).
//   * the WORKER CTAs (ranks 1..C-1) stream the far part of the factors.  A warp owns two 16-row blocks from their
//     birth (load of b) through all far panels (distance kl/16 .. GC_D+1 blocks from the diagonal), one row per lane,
//     16 FMAs per panel, and then hands the values to the leader.  The next panel's factor entries are loaded before the
//     wait for the current panel's solution (register double buffer), so they are in flight while the chain is busy.
// Per element the operations and their order are exactly those of DGBTRS 'N' (SURVEY.md A.4): forward
// b[i] = fma(-b[j], L[i,j], b[i]) for j ascending; backward b[j] = b[j] / U[j,j] (true quotient, gb_div), then
// b[i] = fma(-b[j], U[i,j], b[i]) for j descending; entries outside the band or the matrix are skipped, not multiplied
// by zero.  The result is therefore bit-identical to the reference CPU path.
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

#define GC_NB 16        // panel width = row-block height
#define GC_D 6          // row blocks next to the diagonal block that the leader updates itself
#define GC_LW 7         // leader warps working on row blocks (>= GC_D + 1)
#define GC_HS 16        // hand-off ring slots (power of two, >= GC_D + 2)
#define GC_THREADS (32 * (GC_LW + 1))
#define GC_SPIN_LIMIT (1u << 26)

// Every value that crosses warps or CTAs travels as a 16-byte cell {lo, tag, hi, tag}: two 8-byte halves that each carry
// the tag, so a reader that sees both tags has the whole double even if the 16-byte store were split (8-byte accesses
// are single-copy atomic).  No flags, no fences: a poll of the cell IS the read of the value.
struct __align__(16) GcCell {
    unsigned lo, tag0, hi, tag1;
};

__device__ __forceinline__ void gc_put(GcCell *p, double v, unsigned tag)  // own shared memory
{
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"((unsigned)__double2loint(v)), "r"(tag),
                 "r"((unsigned)__double2hiint(v)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ void gc_put_remote(GcCell *own, unsigned rank, double v, unsigned tag)  // peer CTA's copy of `own`
{
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((unsigned)__cvta_generic_to_shared(own)), "r"(rank));
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "r"((unsigned)__double2loint(v)), "r"(tag),
                 "r"((unsigned)__double2hiint(v)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ bool gc_get(const GcCell *p, unsigned want, double &v)  // own shared memory
{
    unsigned lo, t0, hi, t1;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1)
                 : "r"((unsigned)__cvta_generic_to_shared(p))
                 : "memory");
    v = __hiloint2double((int)hi, (int)lo);
    return t0 == want && t1 == want;
}
__device__ __forceinline__ unsigned gc_tag(const GcCell *p)
{
    unsigned t;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(t) : "r"((unsigned)__cvta_generic_to_shared(&p->tag1)) : "memory");
    return t;
}

struct GcAbort {
    int *g;
    bool dead;
    unsigned spins;
    __device__ __forceinline__ bool tick()  // called once per failed poll; true when the wait must be given up
    {
        if ((++spins & 0x3ffu) == 0) {
            if (*(volatile int *)g) dead = true;
            else if (spins >= GC_SPIN_LIMIT) { atomicExch(g, 1); dead = true; }
        }
        return dead;
    }
};

// all 16 cells of a panel (every lane reads the same addresses: broadcast).  Waiting polls ONE tag word -- a waiting warp
// must not flood the shared-memory pipe that the chain warp's shuffles go through -- then reads and verifies the panel.
__device__ __forceinline__ void gc_read_panel(const GcCell *slot, unsigned want, double (&xv)[GC_NB], GcAbort &ab)
{
    if (ab.dead) return;
    ab.spins = 0;
    for (;;) {
        if (gc_tag(slot + (GC_NB - 1)) == want) {
            bool ok = true;
#pragma unroll
            for (int c = 0; c < GC_NB; ++c) ok &= gc_get(slot + c, want, xv[c]);
            if (ok) return;
        }
        if (ab.tick()) return;
    }
}
// leader-local panel: poll the tag, then 8 x LDS.128
__device__ __forceinline__ void gc_read_local(const double *row, const unsigned *tagp, unsigned want, double (&xv)[GC_NB], GcAbort &ab)
{
    if (ab.dead) return;
    ab.spins = 0;
    const unsigned ta = (unsigned)__cvta_generic_to_shared(tagp);
    for (;;) {
        unsigned tg;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tg) : "r"(ta) : "memory");
        if (tg == want) break;
        if (ab.tick()) return;
    }
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    const unsigned ra = (unsigned)__cvta_generic_to_shared(row);
#pragma unroll
    for (int c = 0; c < GC_NB; c += 2)
        asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(xv[c]), "=d"(xv[c + 1]) : "r"(ra + 8u * c) : "memory");
}
__device__ __forceinline__ void gc_wait_local(const unsigned *tagp, unsigned want, GcAbort &ab)  // tags only grow
{
    if (ab.dead) return;
    ab.spins = 0;
    const unsigned ta = (unsigned)__cvta_generic_to_shared(tagp);
    for (;;) {
        unsigned tg;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tg) : "r"(ta) : "memory");
        if ((int)(tg - want) >= 0 || ab.tick()) return;
    }
}
__device__ __forceinline__ void gc_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gc_mbar_arrive(unsigned long long *bar)
{
    unsigned long long st;
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// leader-local panel through the slot's mbarrier: the waiting warp is suspended by the hardware instead of spinning
__device__ __forceinline__ void gc_read_local_mbar(const double *row, unsigned long long *bar, unsigned parity, double (&xv)[GC_NB], GcAbort &ab)
{
    if (ab.dead) return;
    ab.spins = 0;
    const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
    for (;;) {
        unsigned done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(ba), "r"(parity) : "memory");
        if (done) break;
        ab.spins += 0x3ffu;  // a failed try_wait has already slept for the hardware time-out
        if (ab.tick()) return;
    }
    const unsigned ra = (unsigned)__cvta_generic_to_shared(row);
#pragma unroll
    for (int c = 0; c < GC_NB; c += 2)
        asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(xv[c]), "=d"(xv[c + 1]) : "r"(ra + 8u * c) : "memory");
}
// wait until the slot's tag has reached `want` (tags only grow within a sweep): flow control, no data consumed
__device__ __forceinline__ void gc_wait_tag(const GcCell *cell, unsigned want, GcAbort &ab)
{
    if (ab.dead) return;
    ab.spins = 0;
    while ((int)(gc_tag(cell) - want) < 0)
        if (ab.tick()) return;
}

__device__ __forceinline__ void gc_prefetch_l2_range(const double *p, int ndoubles)
{
    if (ndoubles <= 0) return;
    const unsigned long long a = (unsigned long long)p & ~15ull;
    const unsigned bytes = (unsigned)((((unsigned long long)(p + ndoubles) + 15ull) & ~15ull) - a);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
}

// Factor entries that panel bs (block distance d >= 1) contributes to the row of lane i: v[c] = A[row, 16*bs + c];
// returns the mask of entries inside the band and the matrix.  Forward: L below the diagonal; backward: U above it.
template <bool FWD, bool STREAM>
__device__ __forceinline__ unsigned gc_load_row(const double *__restrict__ ab, i64 ldab, int kv, int bw, i64 n, bool active, i64 bs, int d,
                                                int i, double (&v)[GC_NB])
{
    const double *p = ab + (bs * GC_NB) * ldab + (FWD ? kv + GC_NB * d + i : kv - GC_NB * d + i);  // entry for c = 0
    const i64 st = ldab - 1;  // next column, same row
    // interior of the band (the common case): 16 independent loads, no masks.  A lone warp issues dependent integer
    // code at 4-6 cycles per instruction, so the per-entry predicate arithmetic below costs more than the loads.
    if (active && GC_NB * d + (GC_NB - 1) <= bw && (FWD || (bs + 1) * GC_NB <= n)) {
#pragma unroll
        for (int c = 0; c < GC_NB; ++c) v[c] = STREAM ? ld_stream(p + c * st) : __ldg(p + c * st);
        return 0xffffu;
    }
    unsigned vm = 0;
#pragma unroll
    for (int c = 0; c < GC_NB; ++c) {
        const int off = FWD ? GC_NB * d + i - c : GC_NB * d + c - i;  // |row - col| >= 1
        const bool ok = active && off <= bw && (FWD || bs * GC_NB + c < n);
        v[c] = ok ? (STREAM ? ld_stream(p + c * st) : __ldg(p + c * st)) : 0.0;
        vm |= (unsigned)ok << c;
    }
    return vm;
}
// p ? fma(a, b, c) : c as ONE predicated DFMA (a select after the FMA would add its latency to every link of the chain)
__device__ __forceinline__ double gc_fma_if(unsigned p, double a, double b, double c)
{
    asm("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q fma.rn.f64 %0, %2, %3, %0; }" : "+d"(c) : "r"(p), "d"(a), "d"(b));
    return c;
}
template <bool FWD>
__device__ __forceinline__ double gc_apply(double acc, const double (&xp)[GC_NB], const double (&v)[GC_NB], unsigned vm)
{
    if (vm == 0xffffu) {  // interior of the band: no masks on the chain
        if (FWD) {
#pragma unroll
            for (int c = 0; c < GC_NB; ++c) acc = fma(-xp[c], v[c], acc);
        } else {
#pragma unroll
            for (int c = GC_NB - 1; c >= 0; --c) acc = fma(-xp[c], v[c], acc);
        }
    } else if (FWD) {
#pragma unroll
        for (int c = 0; c < GC_NB; ++c) acc = gc_fma_if((vm >> c) & 1u, -xp[c], v[c], acc);
    } else {
#pragma unroll
        for (int c = GC_NB - 1; c >= 0; --c) acc = gc_fma_if((vm >> c) & 1u, -xp[c], v[c], acc);
    }
    return acc;
}

// Leader-local copy of the newest panels as plain doubles + one tag per panel (data, CTA fence, tag): the warp that
// continues the chain reads a panel with 8 LDS.128 instead of 16 tagged cells -- on a lone warp every instruction of the
// hand-over costs 3-5 cycles of chain time.
struct GcLocal {
    double xl[GC_HS][GC_NB];
    unsigned long long lbar[GC_HS];  // one mbarrier per ring slot: the warp that continues the chain SLEEPS on it (try_wait)
    unsigned ltag[GC_HS];            // same information as a tag, for the prefetch warp (may lag more than a ring)
    long long pubclk[16];  // BMB200_GBTRS_STATS only
};

// One sweep.  Step t handles row block blk(t) (forward: t, backward: PB-1-t); the triangle of step t needs the updates
// of steps t-KB .. t-1 applied to its rows in that order: far steps [t-KB, t-DN-1] by a worker, near steps by the leader.
// Shared memory of every CTA: hand[GC_HS][16] cells (used in the leader), then xs[xslots][16] cells.
// `kv` = row of the diagonal inside a column of `ab`, `bw` = reach of the sweep below (FWD) / above (!FWD) the diagonal,
// DIV = divide by the diagonal (non-unit triangle): gbtrs is <FWD,no DIV>(kv=kl+ku, bw=kl) then <!FWD,DIV>(kv, bw=kv);
// tbsv 'U' is <!FWD, DIV = non-unit>(kv=k, bw=k); tbsv 'L' is <FWD, DIV = non-unit>(kv=0, bw=k).
template <bool FWD, bool DIV>
__device__ __forceinline__ void gc_sweep(cg::cluster_group &cluster, GcCell *smem, int xslots, i64 n, int kv, int bw,
                                         const double *__restrict__ ab, i64 ldab, double *x, int *gabort, int pfdist, long long *stats, GcLocal &loc,
                                         bool after_fwd = false)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, i = lane & 15;
    const i64 PB = (n + GC_NB - 1) / GC_NB;
    const int KB = (bw + GC_NB - 1) / GC_NB;        // largest block distance with an in-band entry
    const int DN = KB < GC_D ? KB : GC_D;
    const unsigned fbase = FWD ? 0u : (unsigned)PB;  // backward tags continue after the forward ones: tags only ever grow
    const int xmask = xslots - 1;
    const unsigned rank = cluster.block_rank(), C = cluster.num_blocks();
    GcCell *hand = smem;                            // [GC_HS][16]
    GcCell *xs = smem + GC_HS * GC_NB;              // [xslots][16]
    GcAbort abt{gabort, false, 0u};
    long long (&pubclk)[16] = loc.pubclk;
    double (&xl)[GC_HS][GC_NB] = loc.xl;
    unsigned (&ltag)[GC_HS] = loc.ltag;
    unsigned long long (&lbar)[GC_HS] = loc.lbar;

    if (rank == 0) {
        if (wid < GC_LW) {
            // ------------------------------------------ leader: chain + near updates ------------------------------
            long long st[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long sw_clk = 0;
            unsigned long long sw_ns = 0;
            if (stats) { sw_clk = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(sw_ns)); }
#define GC_TICK(var) long long var = 0; if (stats) { asm volatile("" : "+d"(xi)); var = clock64(); }
            for (i64 t = wid; t < PB; t += GC_LW) {
                const i64 b = FWD ? t : PB - 1 - t;
                const i64 r = b * GC_NB + i;
                const bool rowok = r < n;
                const i64 sN = (t > DN) ? t - DN : 0;  // near steps are [sN, t)
                long long cs = 0;
                if (stats) cs = clock64();
                double va[GC_NB], vb[GC_NB];
                unsigned ma = 0, mb = 0;
                if (sN < t) ma = gc_load_row<FWD, false>(ab, ldab, kv, bw, n, rowok, FWD ? sN : PB - 1 - sN, (int)(t - sN), i, va);
                // operands of the diagonal triangle (independent of the right-hand side: loaded before any wait)
                double T[GC_NB];
                unsigned tm = 0;
                const bool tint = rowok && (b + 1) * GC_NB <= n && bw >= GC_NB - 1;  // whole 16 x 16 block in band and matrix
                if (tint) {  // unmasked loads (the other triangle's entries are in bounds and discarded below)
                    const double *pt = ab + (kv + i) + (b * GC_NB) * ldab;
#pragma unroll
                    for (int c = 0; c < GC_NB; ++c) T[c] = __ldg(pt + c * (ldab - 1));
                    tm = FWD ? ((1u << i) - 1u) : (0xffffu & ~((2u << i) - 1u));
                } else {
#pragma unroll
                    for (int c = 0; c < GC_NB; ++c) {
                        const i64 col = b * GC_NB + c;
                        const bool ok = FWD ? (rowok && c < i && i - c <= bw) : (c > i && col < n && c - i <= bw);
                        T[c] = ok ? __ldg(ab + (kv + i - c) + col * ldab) : 0.0;
                        tm |= (unsigned)ok << c;
                    }
                }
                double Ud = 1.0, rcp = 1.0;
                bool dsafe = false;
                if (DIV) {
                    Ud = rowok ? __ldg(ab + kv + r * ldab) : 1.0;
                    rcp = 1.0 / Ud;  // one IEEE reciprocal per row, off the chain (gb_div: two Markstein corrections)
                    dsafe = gb_div_safe_divisor(Ud);
                }
                const unsigned mytag = (unsigned)(t + 1) + fbase;
                double xi = 0.0;
                GC_TICK(c0)
                {   // hand-off from the worker that carried this block through its far panels
                    const GcCell *hc = hand + (int)(t & (GC_HS - 1)) * GC_NB + i;
                    abt.spins = 0;
                    while (!abt.dead && !gc_get(hc, mytag, xi))
                        if (abt.tick()) break;
                }
                GC_TICK(c1)
                // near steps, the next step's factor entries in flight while this step waits for its panel
                auto step = [&](i64 s, const double (&vc)[GC_NB], unsigned mc, double (&vn)[GC_NB], unsigned &mn) {
                    GC_TICK(cl)
                    if (s + 1 < t) mn = gc_load_row<FWD, false>(ab, ldab, kv, bw, n, rowok, FWD ? s + 1 : PB - 2 - s, (int)(t - s - 1), i, vn);
                    double xp[GC_NB];
                    GC_TICK(ca)
                    st[9] += ca - cl;
                    {
                        // phase of slot (s mod 16) for panel s: its uses so far = s/16 in this sweep, plus -- when a forward sweep
                        // ran first on the same barriers -- one per forward panel that mapped to the slot
                        const int sl16 = (int)(s & (GC_HS - 1));
                        const i64 prev = (after_fwd && sl16 < PB) ? (PB - sl16 + GC_HS - 1) / GC_HS : 0;
                        gc_read_local_mbar(xl[sl16], &lbar[sl16], (unsigned)((s / GC_HS + prev) & 1), xp, abt);
                    }
                    GC_TICK(cb)
                    xi = gc_apply<FWD>(xi, xp, vc, mc);
                    GC_TICK(cc)
                    if (s + 1 == t) { st[1] += cb - ca; st[2] += cc - cb; if (stats) st[3] += cb - *(volatile long long *)&pubclk[s & 15]; } else { st[4] += cc - cb; }
                };
                for (i64 s = sN; s < t; s += 2) {
                    step(s, va, ma, vb, mb);
                    if (s + 1 < t) step(s + 1, vb, mb, va, ma);
                }
                GC_TICK(c2)
                // 16 x 16 triangle: both half-warps run the same chain (lanes 16..31 mirror 0..15)
                // Wide band (>= 15): the only masked entries are those of rows that are already final (T[c] = 0 there), so
                // the chain runs WITHOUT selects -- each lane banks its value when its own column comes up ("fin") and
                // whatever a zero multiplier does to it afterwards is never looked at.  Rows/columns outside the matrix
                // carry exact zeros (hand-off 0, T = 0, Ud = 1), which leave every valid row bit-unchanged.
                if (tint) {
#pragma unroll
                    for (int c = 0; c < GC_NB; ++c) T[c] = ((tm >> c) & 1u) ? T[c] : 0.0;
                }
                double fin = xi;
                if (!DIV) {
                    if (bw >= GC_NB - 1) {
#pragma unroll
                        for (int cc = 0; cc < GC_NB - 1; ++cc) {  // the last column has no row left to update
                            const int c = FWD ? cc : GC_NB - 1 - cc;
                            const double u = __shfl_sync(0xffffffffu, xi, c);
                            fin = (i == c) ? xi : fin;
                            xi = fma(-u, T[c], xi);
                        }
                        xi = (i == (FWD ? GC_NB - 1 : 0)) ? xi : fin;
                    } else {
#pragma unroll
                        for (int cc = 0; cc < GC_NB; ++cc) {
                            const int c = FWD ? cc : GC_NB - 1 - cc;
                            const double u = __shfl_sync(0xffffffffu, xi, c);
                            xi = gc_fma_if((tm >> c) & 1u, -u, T[c], xi);
                        }
                    }
                } else {
                    // Speculative pass: every quotient by two Markstein corrections of x * RN(1/d) (= the IEEE quotient
                    // whenever the operands are in the safe range, see gb_div); every lane forms the quotient of its own
                    // row and lane c's is the one broadcast, so the warp stays converged and no call sits inside the chain.
                    const double xsave = xi;
                    bool slow = false;
                    if (bw >= GC_NB - 1) {
#pragma unroll
                        for (int cc = 0; cc < GC_NB; ++cc) {
                            const int c = FWD ? cc : GC_NB - 1 - cc;
                            const double q0 = __dmul_rn(xi, rcp);
                            const double q1 = fma(fma(-q0, Ud, xi), rcp, q0);
                            const double q2 = fma(fma(-q1, Ud, xi), rcp, q1);
                            const bool mine = (i == c);
                            slow |= mine && rowok && !(dsafe && gb_exp_mid(xi));
                            fin = mine ? q2 : fin;
                            const double q = __shfl_sync(0xffffffffu, q2, c);
                            xi = fma(-q, T[c], xi);
                        }
                        xi = fin;
                    } else {
#pragma unroll
                        for (int cc = 0; cc < GC_NB; ++cc) {
                            const int c = FWD ? cc : GC_NB - 1 - cc;
                            const double q0 = __dmul_rn(xi, rcp);
                            const double q1 = fma(fma(-q0, Ud, xi), rcp, q0);
                            const double q2 = fma(fma(-q1, Ud, xi), rcp, q1);
                            const bool mine = (i == c) && rowok;
                            slow |= mine && !(dsafe && gb_exp_mid(xi));
                            xi = mine ? q2 : xi;
                            const double q = __shfl_sync(0xffffffffu, xi, c);
                            xi = gc_fma_if((tm >> c) & 1u, -q, T[c], xi);
                        }
                    }
                    if (__any_sync(0xffffffffu, slow)) {  // an operand outside the safe range: redo with IEEE divisions
                        xi = xsave;
                        for (int cc = 0; cc < GC_NB; ++cc) {
                            const int c = FWD ? cc : GC_NB - 1 - cc;
                            if (i == c && rowok) xi = gb_div_ieee(xi, Ud);
                            const double q = __shfl_sync(0xffffffffu, xi, c);
                            if ((tm >> c) & 1u) xi = fma(-q, T[c], xi);
                        }
                    }
                }
                GC_TICK(c3)
                // publish: own shared memory first (the next leader warp is polling it) ...
                GcCell *mine = xs + (int)(t & xmask) * GC_NB + i;
                if (stats && lane == 0) *(volatile long long *)&pubclk[t & 15] = c3;
                if (lane < GC_NB) xl[t & (GC_HS - 1)][i] = xi;
                __syncwarp();
                if (lane == 0) {
                    gc_mbar_arrive(&lbar[t & (GC_HS - 1)]);  // release: the 16 stores above are visible to whoever the phase change wakes
                    *(volatile unsigned *)&ltag[t & (GC_HS - 1)] = mytag;
                }
                if (lane < GC_NB && rowok) x[r] = xi;
                // ... then every worker's ring (the two half-warps share the peers)
                for (unsigned w = 1 + (lane >> 4); w < C; w += 2) gc_put_remote(mine, w, xi, mytag);
                GC_TICK(c4)
                st[8] += c0 - cs; st[0] += c1 - c0; st[5] += c3 - c2; st[6] += c4 - c3; st[7] += 1;
            }
            if (stats && wid == 0 && lane == 0 && blockIdx.x == 0) {
                for (int k = 0; k < 8; ++k) stats[(FWD ? 0 : 8) + k] = st[k];
                stats[20 + (FWD ? 0 : 2)] = st[8];
                stats[21 + (FWD ? 0 : 2)] = st[9];
                unsigned long long e_ns;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(e_ns));
                stats[16 + (FWD ? 0 : 2)] = clock64() - sw_clk;
                stats[17 + (FWD ? 0 : 2)] = (long long)(e_ns - sw_ns);
            }
#undef GC_TICK
        } else if (wid == GC_LW && pfdist > 0) {
            // ------------------------------------------ leader: L2 prefetch of the factor panels ahead ---------------
            for (i64 s = 0; s + pfdist < PB; ++s) {
                if (s > 0) gc_wait_local(&ltag[(s - 1) & (GC_HS - 1)], (unsigned)s + fbase, abt);
                if (abt.dead) break;
                const i64 bs = FWD ? s + pfdist : PB - 1 - (s + pfdist);
                const i64 col = bs * GC_NB + lane;
                if (lane < GC_NB && col < n) {
                    if (FWD) {
                        const i64 lm = (n - 1 - col < bw) ? n - 1 - col : bw;
                        gc_prefetch_l2_range(ab + col * ldab + (kv + 1), (int)lm);
                    } else {
                        gc_prefetch_l2_range(ab + col * ldab + (kv - bw), bw + 1);
                    }
                }
            }
        }
    } else {
        // ---------------------------------------------- workers: far updates --------------------------------------
        // A warp owns 32 consecutive rows = steps t = 2TT (lanes 0..15) and 2TT+1 (lanes 16..31) and walks the panels
        // s = 2TT-KB .. 2TT-DN with all 32 lanes on the same panel (no divergent waits); each half hands its block
        // to the leader after its own last far panel t-DN-1.
        const int NWK = (int)C - 1, widx = (int)rank - 1;
        const int nwarp = GC_THREADS / 32;
        for (i64 TT = widx + (i64)NWK * wid; 2 * TT < PB; TT += (i64)NWK * nwarp) {
            const i64 t = 2 * TT + (lane >> 4);
            const bool tok = t < PB;
            const i64 b = FWD ? t : PB - 1 - t;
            const i64 r = b * GC_NB + i;
            const bool rowok = tok && r < n;
            double bv = rowok ? __ldcg(x + r) : 0.0;  // L2: the forward sweep's results were written by the leader's SM
            const i64 sA = (2 * TT > KB) ? 2 * TT - KB : 0, sE = 2 * TT + 1 - DN;  // panels [sA, sE); lane active while s < t-DN
            bool handed = !tok;
            auto handoff = [&]() {
                // the hand-off slot of block t-GC_HS is free once that block's solution has been published
                if (t >= GC_HS) gc_wait_tag(xs + (int)((t - GC_HS) & xmask) * GC_NB, (unsigned)(t - GC_HS + 1) + fbase, abt);
                gc_put_remote(hand + (int)(t & (GC_HS - 1)) * GC_NB + i, 0u, bv, (unsigned)(t + 1) + fbase);
                handed = true;
            };
            double va[GC_NB], vb[GC_NB];
            unsigned ma = 0, mb = 0;
            if (sA < sE) ma = gc_load_row<FWD, true>(ab, ldab, kv, bw, n, rowok && sA < t - DN, FWD ? sA : PB - 1 - sA, (int)(t - sA), i, va);
            auto step = [&](i64 s, const double (&vc)[GC_NB], unsigned mc, double (&vn)[GC_NB], unsigned &mn) {
                if (s + 1 < sE) mn = gc_load_row<FWD, true>(ab, ldab, kv, bw, n, rowok && s + 1 < t - DN, FWD ? s + 1 : PB - 2 - s, (int)(t - s - 1), i, vn);
                if (!handed && s >= t - DN) handoff();
                double xp[GC_NB];
                gc_read_panel(xs + (int)(s & xmask) * GC_NB, (unsigned)(s + 1) + fbase, xp, abt);
                bv = gc_apply<FWD>(bv, xp, vc, mc);
            };
            for (i64 s = sA; s < sE; s += 2) {
                step(s, va, ma, vb, mb);
                if (s + 1 < sE) step(s + 1, vb, mb, va, ma);
            }
            if (!handed) handoff();
        }
    }
}

// MODE 0: gbtrs (unit forward sweep with the multipliers below row kl+ku, then dividing backward sweep over kl+ku rows);
// MODE 1: backward sweep only, unit diagonal (tbsv 'U','N','U'); MODE 2 / 3: forward sweep only, unit / non-unit
// diagonal in row 0 of the band array (tbsv 'L').  tbsv 'U','N','N' is MODE 0 with kl = 0.
template <int MODE>
__global__ void __launch_bounds__(GC_THREADS, 1)
gbtrs_cluster_noswap(i64 n, int kl, int ku, const double *__restrict__ ab, i64 ldab, double *__restrict__ bmat, i64 ldb, int xslots,
                     int *gabort, int pfdist, long long *stats)
{
    extern __shared__ __align__(16) unsigned char gc_smem_raw[];
    GcCell *gc_smem = reinterpret_cast<GcCell *>(gc_smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned C = cluster.num_blocks();
    double *x = bmat + (i64)(blockIdx.x / C) * ldb;
    __shared__ __align__(16) GcLocal loc;
    for (int k = threadIdx.x; k < (GC_HS + xslots) * GC_NB; k += blockDim.x) gc_smem[k] = GcCell{0u, 0u, 0u, 0u};
    if (threadIdx.x < GC_HS) { loc.ltag[threadIdx.x] = 0u; gc_mbar_init(&loc.lbar[threadIdx.x], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    cluster.sync();
    if (MODE == 0) {
        if (kl > 0) gc_sweep<true, false>(cluster, gc_smem, xslots, n, kl + ku, kl, ab, ldab, x, gabort, pfdist, stats, loc);
        cluster.sync();  // forward results are in global memory; every ring is quiescent
        gc_sweep<false, true>(cluster, gc_smem, xslots, n, kl + ku, kl + ku, ab, ldab, x, gabort, pfdist, stats, loc, kl > 0);
    } else if (MODE == 1) {
        gc_sweep<false, false>(cluster, gc_smem, xslots, n, ku, ku, ab, ldab, x, gabort, pfdist, stats, loc);
    } else if (MODE == 2) {
        gc_sweep<true, false>(cluster, gc_smem, xslots, n, 0, kl, ab, ldab, x, gabort, pfdist, stats, loc);
    } else {
        gc_sweep<true, true>(cluster, gc_smem, xslots, n, 0, kl, ab, ldab, x, gabort, pfdist, stats, loc);
    }
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

// returns 1 when not applicable (the caller then runs the single-CTA kernel), 0 on success, <0 on error.
// mode 0: the caller has verified ipiv = 1:n.  modes 1-3: triangular band solves (see the kernel).
int bmb_cluster_solve(bmb200_ctx *h, int mode, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb)
{
    const int csize_env = h->tune.gbtrs_cluster, pfdist = h->tune.gbtrs_pfdist;
    if (h->tune.gbtrs_nocluster) return 1;
    const i64 kv = kl + ku;
    if (n >= ((i64)1 << 33)) return 1;  // flag arithmetic is 32-bit
    const int KBmax = (int)((kv + GC_NB - 1) / GC_NB);
    int xslots = 32;
    while (xslots < KBmax + GC_D + 4) xslots <<= 1;
    const size_t smem = (size_t)(GC_HS + xslots) * GC_NB * sizeof(GcCell);
    if (smem > 200 * 1024) return 1;
    typedef void (*kern_t)(i64, int, int, const double *, i64, double *, i64, int, int *, int, long long *);
    static const kern_t kerns[4] = {gbtrs_cluster_noswap<0>, gbtrs_cluster_noswap<1>, gbtrs_cluster_noswap<2>, gbtrs_cluster_noswap<3>};
    if (mode < 0 || mode > 3) return 1;
    const kern_t kern = kerns[mode];
    BMB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BMB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int *gabort = h->d_info + 17;
    BMB_CUDA(h, cudaMemsetAsync(gabort, 0, sizeof(int), h->stream));
    // cluster size 8 (portable; 15 clusters co-resident): measured equal to 16 at l=u=1024 -- the leader chain is the bound
    int csize = csize_env > 1 ? csize_env : 8;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    for (;; csize >>= 1) {
        if (csize < 2) return 1;
        cfg.gridDim = dim3((unsigned)(nrhs * csize));
        cfg.blockDim = dim3(GC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = h->stream;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)csize;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        const cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
        const bool dbg = h->tune.debug != 0;
        if (dbg) fprintf(stderr, "[bmb200] gbtrs_cluster: cluster size %d -> %s, %d co-resident clusters, smem %zu\n", csize, cudaGetErrorString(e), nclusters, smem);
        if (e == cudaSuccess && nclusters > 0) break;
        (void)cudaGetLastError();
    }
    const bool want_stats = h->tune.gbtrs_stats != 0;  // development aid: cycle breakdown of leader warp 0
    long long *dstats = nullptr;
    if (want_stats) {
        if (bmb_ensure_scratch(h, 24 * sizeof(long long)) != 0) return BMB200_ERR_CUDA;
        dstats = (long long *)h->scratch;
        BMB_CUDA(h, cudaMemsetAsync(dstats, 0, 24 * sizeof(long long), h->stream));
    }
    BMB_CUDA(h, cudaLaunchKernelEx(&cfg, kern, n, (int)kl, (int)ku, dAB, ldab, dB, ldb, xslots, gabort, pfdist, dstats));
    h->launches++;
    if (want_stats) {
        long long hs[24];
        BMB_CUDA(h, cudaMemcpyAsync(hs, dstats, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
        BMB_CUDA(h, cudaStreamSynchronize(h->stream));
        fprintf(stderr, "[bmb200] gbtrs_cluster sweeps: fwd %lld cycles in %lld ns, bwd %lld cycles in %lld ns\n", hs[16], hs[17], hs[18], hs[19]);
        for (int d = 0; d < 2; ++d) {
            const long long *q = hs + 8 * d, nb = q[7] ? q[7] : 1;
            fprintf(stderr, "[bmb200] gbtrs_cluster %s, leader warp 0, cycles per block (%lld blocks): hand-off wait %lld | last near step: wait %lld apply %lld | "
                            "publish->detect latency %lld, earlier near steps apply %lld | triangle %lld | publish %lld | setup %lld | near loads %lld\n",
                    d ? "bwd" : "fwd", q[7], q[0] / nb, q[1] / nb, q[2] / nb, q[3] / nb, q[4] / nb, q[5] / nb, q[6] / nb, hs[20 + 2 * d] / nb, hs[21 + 2 * d] / nb);
        }
    }
    int ha = 0;
    BMB_CUDA(h, cudaMemcpyAsync(&ha, gabort, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BMB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (ha) {
        snprintf(h->err, sizeof(h->err), "dgbtrs: cluster pipeline timed out waiting for a peer CTA");
        return BMB200_ERR_CUDA;
    }
    return 0;
}

int bmb_gbtrs_cluster(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, double *dB, i64 ldb)
{
    return bmb_cluster_solve(h, 0, n, kl, ku, nrhs, dAB, ldab, dB, ldb);
}
