// gbtrs_slot.cu -- multi-RHS band solve (DGBTRS 'N') for narrow bands with interchanges: "slot-scheduled" sweeps.
//
// Same arithmetic contract as gbtrs.cu (DGBTRS 'N'; reference call site src/banded/linalg.jl:28): forward =
// row interchange, then b[j+k] = fma(-b[j], L[k,j], b[j+k]); backward = TRUE division by the diagonal, then
// b[j-k] = fma(-x[j], U[j-k,j], b[j-k]).  Every value receives exactly the reference's operations in the reference's
// order, so the solution is bit-identical; only WHERE a value lives and WHEN an operation is issued change.
//
// A sweep is a chain of n dependent steps per right-hand side.  The older kernels pay one warp shuffle + one FMA (and
// one IEEE division) of latency per step (gbtrs_shfl.cu: 173 cycles per forward+backward step at C4).  Here:
//   * one warp = one right-hand side; a window value stays in ONE lane ("slot") from the step it enters the window to the
//     step it becomes a solution entry.  The row interchanges of the factorisation only re-label which value is the
//     pivot row -- they never move data.  Which slot holds the pivot row of every step, and which multiplier every slot
//     needs at every step, depend only on (AB, ipiv): a prepass kernel turns them into a SCHEDULE stream (one record per
//     block of P steps: P pivot slots, a P x P triangle of multipliers, P multipliers per slot, presence masks).
//   * the sweep advances P steps per round: P independent shuffles broadcast the P pivot values, every lane solves the
//     P x P triangle redundantly (its chain is P-1 dependent FMAs in registers), then applies its own P updates.  The
//     cross-lane latency (26 cycles) is paid once per P steps instead of once per step.
//   * backward sweep: the reciprocal of every diagonal entry is formed in the prepass as a two-term value
//     r_hi + r_lo; q = fma(t, r_hi, t*r_lo) is within 2^-104 of t/d before its final rounding, and one Markstein
//     correction q2 = fma(fma(-q,d,t), r_hi, q) is RN(t/d) (r_hi correctly rounded, q faithful).  The chain continues with
//     q (two dependent operations); q2 == q is verified off the chain and a block where the test fails (probability
//     ~2^-51 per division; also zero / tiny / huge / non-finite operands) is redone with the IEEE division.
//   * the schedule stream is fetched per CTA by TMA bulk copies (cp.async.bulk + mbarrier full/empty ring, one producer
//     warp) and shared by the CTA's warps; B moves through a per-warp shared-memory ring in coalesced 256-byte chunks
//     (cp.async prefetch, one store per 32 finished rows).
#include "common.cuh"

#define SL_FULL 0xffffffffu
#define SL_COLS 64   // columns per schedule stage (= renormalisation chunk of the forward sweep)
#define SL_NS 4      // stages in the shared-memory ring
#define SL_RB 512    // B ring rows per right-hand side (power of two)
#define SL_LA 256    // B rows prefetched ahead (multiple of 32)

template <int P>
struct SlotFmt {
    static constexpr int NTRI = P * (P - 1) / 2;
    static constexpr int NT = ((NTRI + 1) / 2) * 2;
    static constexpr int OFF_M = 0;                   // double2 M2[P/2][32]: NEGATED multipliers of slot `lane` for steps 2c2, 2c2+1
    static constexpr int OFF_T = 256 * P;             // double T[NT]: T[i(i-1)/2 + c] = NEGATED multiplier applied at step c to the pivot of step i
    static constexpr int OFF_D = OFF_T + 8 * NT;      // double D[P], RH[P], RL[P]: diagonal, reciprocal hi / lo (backward)
    static constexpr int OFF_RH = OFF_D + 8 * P;
    static constexpr int OFF_RL = OFF_RH + 8 * P;
    static constexpr int OFF_CODE = OFF_RL + 8 * P;   // u32 code[32]: bits 0..7 own-update mask; bits 8..15: 8*(c+1) when this lane holds the
                                                      // pivot row of step c (0 otherwise); bit 31: a row enters through this lane
                                                      // after the shuffles, bits 16..27 = its byte offset in the B ring ((row mod SL_RB) * 8)
    static constexpr int OFF_SRC = OFF_CODE + 128;    // i32 src[P]: pivot slot of step c (16-byte aligned, padded to 16 bytes)
    static constexpr int OFF_TM = OFF_SRC + (4 * P < 16 ? 16 : 4 * P);  // u32 tmask (bit i(i-1)/2 + c), 12 bytes pad
    static constexpr int BLK = OFF_TM + 16;
    static constexpr unsigned TFULL = (NTRI >= 32) ? 0xffffffffu : ((1u << NTRI) - 1u);
    static constexpr int G = SL_COLS / P;             // blocks per stage
    static constexpr int OFF_END = G * BLK;           // u8 endmap[32]: forward renormalisation at the end of the stage
    static constexpr int STAGE = OFF_END + 32;
    static_assert(BLK % 16 == 0 && STAGE % 16 == 0 && OFF_T % 16 == 0 && OFF_D % 16 == 0, "bulk-copy / vector alignment");
};

// ---------------------------------------------------------------------------------------------------------------------
// prepass, forward sweep: one warp per stage (64 columns), lanes = slots.  Canonical state at a stage start: rows
// j0 .. j0+kl+P-1 (the kl carried window rows and the P rows that enter during the first block) sit in slots
// 0 .. kl+P-1; the rows entering during block g+1 are loaded at the end of block g into the slots block g's pivots
// freed.  The solve kernel re-establishes the canonical state at the end of every stage with one shuffle (endmap),
// which is what makes the stages independent here.
// ---------------------------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(128)
slot_build_fwd(i64 n, int kl, int kv, const double *__restrict__ ab, i64 ldab, const i64 *__restrict__ ipiv,
               unsigned char *__restrict__ sched, i64 nstages)
{
    using F = SlotFmt<P>;
    const int lane = threadIdx.x & 31;
    const i64 s = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= nstages) return;
    unsigned char *st = sched + s * F::STAGE;
    const i64 j0 = s * SL_COLS;
    int pv[SL_COLS / 32];  // pivot rows of the stage relative to j0
#pragma unroll
    for (int t = 0; t < SL_COLS / 32; ++t) {
        const i64 j = j0 + 32 * t + lane;
        pv[t] = (j < n) ? (int)(ipiv[j] - 1 - j0) : -1;
    }
    int pos = (lane < kl + P && j0 + lane < n) ? lane : -1;  // row (relative to j0) held by this slot, -1 = free
    for (int g = 0; g < F::G; ++g) {
        const int jrel = g * P;
        const i64 jb = j0 + jrel;
        unsigned char *blk = st + g * F::BLK;
        unsigned om = 0, ent = 0u;
        double Mr[P];
        int src[P];
        bool real[P];
#pragma unroll
        for (int c = 0; c < P; ++c) {
            Mr[c] = 0.0;
            src[c] = 0;
            const i64 j = jb + c;
            const int jr = jrel + c;
            real[c] = j < n;
            if (real[c]) {
                const int piv = __shfl_sync(SL_FULL, pv[jr >> 5], jr & 31);
                const unsigned ms = __ballot_sync(SL_FULL, pos == piv), mf = __ballot_sync(SL_FULL, pos == jr);
                const int f = mf ? __ffs(mf) - 1 : 0;
                const int sl = ms ? __ffs(ms) - 1 : f;  // an out-of-range pivot index means "no interchange"
                src[c] = sl;
                if (sl != f && lane == f) pos = piv;
                if (lane == sl) pos = -1;
                const i64 rem = n - 1 - j;
                const int km = rem < kl ? (int)rem : kl;
                if (pos > jr && pos <= jr + km) {
                    Mr[c] = -ab[(i64)(kv + pos - jr) + j * ldab];  // -L[pos-jr-1, j] = -ab[kv+1+(pos-jr-1), j]
                    om |= 1u << c;
                }
            }
        }
        unsigned tmask = 0;
#pragma unroll
        for (int i = 1; i < P; ++i)
#pragma unroll
            for (int c = 0; c < i; ++c) {
                const int idx = i * (i - 1) / 2 + c;
                const double t = __shfl_sync(SL_FULL, Mr[c], src[i]);
                const unsigned o = __shfl_sync(SL_FULL, om, src[i]);
                const bool present = real[i] && ((o >> c) & 1u);
                if (present) tmask |= 1u << idx;
                if (lane == 0) reinterpret_cast<double *>(blk + F::OFF_T)[idx] = present ? t : 0.0;
            }
        // a pivot lane's register is dead once its value has been broadcast: it takes the row that enters the window
        // at the matching step of the NEXT block (loaded by the sweep right after this block's shuffles)
#pragma unroll
        for (int c = 0; c < P; ++c)
            if (real[c] && lane == src[c]) {
                om = 0;
                ent |= (unsigned)(8 * (c + 1)) << 8;
#pragma unroll
                for (int cc = 0; cc < P; ++cc) Mr[cc] = 0.0;
                if (jb + P + kl + c < n) {
                    pos = jrel + P + kl + c;
                    ent |= 0x80000000u | ((unsigned)((jb + P + kl + c) & (SL_RB - 1)) << 19);  // (row mod SL_RB) * 8 at bit 16
                }
            }
#pragma unroll
        for (int c2 = 0; c2 < P / 2; ++c2)
            reinterpret_cast<double2 *>(blk + F::OFF_M)[c2 * 32 + lane] = make_double2(Mr[2 * c2], Mr[2 * c2 + 1]);
        reinterpret_cast<unsigned *>(blk + F::OFF_CODE)[lane] = om | ent;
        if (lane < P) {
            reinterpret_cast<double *>(blk + F::OFF_D)[lane] = 1.0;
            reinterpret_cast<double *>(blk + F::OFF_RH)[lane] = 1.0;
            reinterpret_cast<double *>(blk + F::OFF_RL)[lane] = 0.0;
        }
        if (lane < P) {
            int sv = 0;
#pragma unroll
            for (int c = 0; c < P; ++c)
                if (lane == c) sv = src[c];
            reinterpret_cast<int *>(blk + F::OFF_SRC)[lane] = sv;
        }
        if (lane == 0) {
            if (F::NT > F::NTRI) reinterpret_cast<double *>(blk + F::OFF_T)[F::NTRI] = 0.0;
            reinterpret_cast<unsigned *>(blk + F::OFF_TM)[0] = tmask;
        }
    }
    unsigned char *em = st + F::OFF_END;
    em[lane] = (unsigned char)lane;
    __syncwarp();
    if (pos >= SL_COLS && pos < SL_COLS + kl + P) em[pos - SL_COLS] = (unsigned char)lane;
}

// ---------------------------------------------------------------------------------------------------------------------
// prepass, backward sweep: no interchanges, so the schedule is closed-form.  Virtual row index v = n-1-row (the sweep
// ascends in v); row v lives in lane v mod 32; the lanes of a block's P pivot rows take the rows v+32 after the shuffle.
// ---------------------------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(128)
slot_build_bwd(i64 n, int kv, const double *__restrict__ ab, i64 ldab, unsigned char *__restrict__ sched, i64 nblocks)
{
    using F = SlotFmt<P>;
    const int lane = threadIdx.x & 31;
    for (i64 bI = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bI < nblocks; bI += (i64)gridDim.x * (blockDim.x >> 5)) {
        const i64 s = bI / F::G;
        const int g = (int)(bI - s * F::G);
        unsigned char *st = sched + s * F::STAGE;
        unsigned char *blk = st + g * F::BLK;
        const i64 vb = bI * P;
        const i64 rowv = vb + P + (i64)((lane - (int)((vb + P) & 31)) & 31);  // the lane's row after the block's entries
        unsigned om = 0, ent = 0u;
        double Mr[P];
#pragma unroll
        for (int c = 0; c < P; ++c) {
            Mr[c] = 0.0;
            const i64 vp = vb + c;
            if (lane == (int)(vp & 31) && vp < n) ent |= (unsigned)(8 * (c + 1)) << 8;
            if (lane == (int)(vp & 31) && vp + 32 < n) ent |= 0x80000000u | ((unsigned)((vp + 32) & (SL_RB - 1)) << 19);
            if (vp < n) {
                const i64 col = n - 1 - vp;
                const i64 dist = rowv - vp;
                if (dist >= 1 && dist <= kv && rowv < n) {
                    Mr[c] = -ab[(i64)(kv - dist) + col * ldab];
                    om |= 1u << c;
                }
            }
        }
        // triangle entry `lane`: (i, c) with lane = i(i-1)/2 + c
        {
            int i = 1, c = lane;
            while (i < P && c >= i) { c -= i; ++i; }
            bool present = false;
            double t = 0.0;
            if (lane < F::NTRI && vb + i < n && (i - c) <= kv) {
                t = -ab[(i64)(kv - (i - c)) + (n - 1 - (vb + c)) * ldab];
                present = true;
            }
            const unsigned tmask = __ballot_sync(SL_FULL, present);
            if (lane < F::NT) reinterpret_cast<double *>(blk + F::OFF_T)[lane] = t;
            if (lane == 0) reinterpret_cast<unsigned *>(blk + F::OFF_TM)[0] = tmask;
        }
        if (lane < P) {
            double d = 1.0, rh = 1.0, rl = 0.0;
            if (vb + lane < n) {
                d = ab[(i64)kv + (n - 1 - (vb + lane)) * ldab];
                if (gb_div_safe_divisor(d)) {
                    rh = 1.0 / d;
                    rl = __dmul_rn(fma(-d, rh, 1.0), rh);
                } else {
                    rh = rl = __longlong_as_double(0x7ff8000000000000ll);  // forces the verified-division test to fail
                }
            }
            reinterpret_cast<double *>(blk + F::OFF_D)[lane] = d;
            reinterpret_cast<double *>(blk + F::OFF_RH)[lane] = rh;
            reinterpret_cast<double *>(blk + F::OFF_RL)[lane] = rl;
        }
#pragma unroll
        for (int c2 = 0; c2 < P / 2; ++c2)
            reinterpret_cast<double2 *>(blk + F::OFF_M)[c2 * 32 + lane] = make_double2(Mr[2 * c2], Mr[2 * c2 + 1]);
        reinterpret_cast<unsigned *>(blk + F::OFF_CODE)[lane] = om | ent;
        if (lane < P) reinterpret_cast<int *>(blk + F::OFF_SRC)[lane] = (int)((vb + lane) & 31);
        if (g == 0) st[F::OFF_END + lane] = (unsigned char)lane;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the sweep
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sl_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sl_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sl_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sl_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sl_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sl_mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sl_smem(bar)) : "memory");
}
__device__ __forceinline__ void sl_mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned a = sl_smem(bar);
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void sl_bulk_g2s(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sl_smem(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(sl_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void sl_cp8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sl_smem(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void sl_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sl_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// explicit 32-bit shared-memory accesses (the schedule ring and the B ring are addressed by hand in the hot loop)
__device__ __forceinline__ double sl_lds64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double2 sl_lds128(unsigned a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned sl_lds32(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 sl_lds128i(unsigned a)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sl_sts128(unsigned a, double v0, double v1)
{
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ void sl_sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sl_sts64_if(unsigned a, double v, int lane, int c)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(a), "d"(v), "r"(lane), "r"(c) : "memory");
}
// acc <- fma(x, m, acc) when bit != 0, acc unchanged otherwise.  An absent update has m == +0.0 in the schedule; with
// x replaced by -0.0 the product is -0 and acc + (-0) == acc for EVERY acc (signed zeros, infinities and NaNs included),
// so the select sits on the operand (off the accumulator's dependency chain) and the FMA itself is unconditional.
__device__ __forceinline__ void sl_fma_if(double &acc, double x, double m, unsigned bit)
{
    // (only the high word is replaced: a negative value with a zero exponent field times +0.0 is -0 all the same)
    const double xe = __hiloint2double(bit ? __double2hiint(x) : (int)0x80000000u, __double2loint(x));
    acc = fma(xe, m, acc);
}

// Schedule words of one block (everything but the pivot slots, which are fetched one block ahead because the shuffles
// need them first): loaded right behind the shuffles, so their latency hides behind the shuffles' own.
template <int P, bool BWD>
struct SlotRegs {
    unsigned tm;
    double T[SlotFmt<P>::NT];
    double D[BWD ? P : 1], RH[BWD ? P : 1], RL[BWD ? P : 1];
};
template <int P, bool BWD>
__device__ __forceinline__ void sl_fetch(SlotRegs<P, BWD> &r, unsigned blk_s)
{
    using F = SlotFmt<P>;
    r.tm = sl_lds32(blk_s + F::OFF_TM);
#pragma unroll
    for (int k = 0; k < F::NT; k += 2) {
        const double2 t2 = sl_lds128(blk_s + F::OFF_T + 8 * k);
        r.T[k] = t2.x, r.T[k + 1] = t2.y;
    }
    if (BWD) {
#pragma unroll
        for (int k = 0; k < P; k += 2) {
            const double2 d2 = sl_lds128(blk_s + F::OFF_D + 8 * k), h2 = sl_lds128(blk_s + F::OFF_RH + 8 * k),
                          l2 = sl_lds128(blk_s + F::OFF_RL + 8 * k);
            r.D[k] = d2.x, r.D[k + 1] = d2.y, r.RH[k] = h2.x, r.RH[k + 1] = h2.y, r.RL[k] = l2.x, r.RL[k + 1] = l2.y;
        }
    }
}
template <int P>
__device__ __forceinline__ void sl_fetch_src(int (&src)[P], unsigned blk_s)
{
    using F = SlotFmt<P>;
    if (P == 2) {
        const int4 a = sl_lds128i(blk_s + F::OFF_SRC);
        src[0] = a.x, src[1] = a.y;
    } else {
#pragma unroll
        for (int c = 0; c < P; c += 4) {
            const int4 a = sl_lds128i(blk_s + F::OFF_SRC + 4 * c);
            src[c] = a.x, src[c + 1] = a.y, src[c + 2] = a.z, src[c + 3] = a.w;
        }
    }
}

// the P x P triangle of one block: xs[i] = (v[i] + sum_{c<i} T[i,c] xs[c]) [/ d_i], T holding the negated multipliers.
// PRED = false is the common case (every triangle entry present).
template <int P, bool BWD, bool PRED>
__device__ __forceinline__ void sl_triangle(const double (&v)[P], double (&xs)[P], const SlotRegs<P, BWD> &r)
{
    unsigned ok = 1u;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        double t = v[i];
#pragma unroll
        for (int c = 0; c < i; ++c) {
            const int idx = i * (i - 1) / 2 + c;
            if (PRED) sl_fma_if(t, xs[c], r.T[idx], r.tm & (1u << idx));
            else t = fma(xs[c], r.T[idx], t);
        }
        if (BWD) {
            const double q = fma(t, r.RH[i], __dmul_rn(t, r.RL[i]));
            const double q2 = fma(fma(-q, r.D[i], t), r.RH[i], q);
            ok &= (unsigned)(q2 == q) & (unsigned)gb_exp_mid(t);
            xs[i] = q;
        } else {
            xs[i] = t;
        }
    }
    if (BWD && !ok) {  // exact redo of the block (rare: see the header)
#pragma unroll
        for (int i = 0; i < P; ++i) {
            double t = v[i];
#pragma unroll
            for (int c = 0; c < i; ++c) {
                const int idx = i * (i - 1) / 2 + c;
                sl_fma_if(t, xs[c], r.T[idx], r.tm & (1u << idx));
            }
            xs[i] = (t == 0.0 && r.RH[i] == r.RH[i]) ? __dmul_rn(t, r.RH[i]) : gb_div_ieee(t, r.D[i]);
        }
    }
}

// R right-hand sides per warp: R independent dependency chains share one instruction stream (and one copy of every
// schedule word), so the issue slots one chain leaves empty while it waits on its own DFMA are taken by the other.
template <int P, bool BWD, int W, int R>
__global__ void __launch_bounds__((W + 1) * 32, 1)
gbtrs_slot(i64 n, int kl, i64 nrhs, const unsigned char *__restrict__ sched, int nstages, double *__restrict__ b, i64 ldb)
{
    using F = SlotFmt<P>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sring = smem;                                                              // SL_NS stages
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sring + SL_NS * F::STAGE);  // [SL_NS]
    unsigned long long *empty = full + SL_NS;                                                 // [SL_NS]
    double *bring_all = reinterpret_cast<double *>(empty + SL_NS);                            // W*R x SL_RB, then W*R pivot buffers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 q0 = (i64)blockIdx.x * (W * R);                                                 // first right-hand side of this CTA
    const i64 left = nrhs - q0;
    const int nact = (int)((left < W * R) ? (left + R - 1) / R : W);                          // warps with at least one right-hand side
    if (threadIdx.x == 0) {
        for (int s = 0; s < SL_NS; ++s) {
            sl_mbar_init(&full[s], 1);
            sl_mbar_init(&empty[s], (unsigned)nact);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == W) {  // ---- producer: streams the schedule through the stage ring ----
        if (lane == 0) {
            for (int s = 0; s < nstages; ++s) {
                const int sl = s % SL_NS;
                if (s >= SL_NS) sl_mbar_wait(&empty[sl], (unsigned)(((s / SL_NS) - 1) & 1));
                sl_mbar_expect_tx(&full[sl], (unsigned)F::STAGE);
                sl_bulk_g2s(sring + sl * F::STAGE, sched + (size_t)s * F::STAGE, (unsigned)F::STAGE, &full[sl]);
            }
        }
        return;
    }
    if (warp >= nact) return;

    const unsigned sring_s = sl_smem(sring);
    auto rowof = [&](i64 v) -> i64 { return BWD ? n - 1 - v : v; };
    double *bring[R];
    unsigned bring_s[R], pbuf_s[R], xprev_s[R];
    double *bcol[R];
    bool live[R];  // a warp's trailing right-hand sides past nrhs shadow its first one (same arithmetic, nothing stored)
    double w[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const i64 q = q0 + (i64)warp * R + r;
        live[r] = q < nrhs;
        bcol[r] = b + (live[r] ? q : q0 + (i64)warp * R) * ldb;
        bring[r] = bring_all + (size_t)(warp * R + r) * SL_RB;
        bring_s[r] = sl_smem(bring[r]);
        // pivot exchange buffer: the lane that holds the pivot row of step c stores it to pbuf[2+c], everyone reads all P
        // back with broadcast LDS.128 (1 + P/2 shared-memory operations instead of 2P shuffles); pbuf[1] is a dummy
        pbuf_s[r] = sl_smem(bring_all + (size_t)W * R * SL_RB) + (unsigned)(warp * R + r) * (8u * P + 16u);
        xprev_s[r] = bring_s[r] + 8u * (SL_RB - P);  // ring address of the previous block's solution entries
        for (int c = 0; c < SL_LA / 32; ++c) {
            const i64 v = 32 * c + lane;
            if (v < n) sl_cp8(bring[r] + v, bcol[r] + rowof(v));
        }
    }
    sl_commit();
    sl_wait<0>();
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) w[r] = ((BWD || lane < kl + P) && lane < n) ? bring[r][lane] : 0.0;
    int slot = 0;
    unsigned phase = 0, vbr8 = 0;  // vbr8 = ring byte offset of the current block's first row
    double xsp[R][P];              // the previous block's solution entries: stored behind the next block's pivot exchange
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < P; ++c) xsp[r][c] = 0.0;
    unsigned code = 0;
    bool tfull = false;
#pragma unroll 1
    for (int s = 0; s < nstages; ++s) {
        sl_mbar_wait(&full[slot], phase);
        const unsigned sp_s = sring_s + (unsigned)slot * F::STAGE;
        code = sl_lds32(sp_s + F::OFF_CODE + 4u * lane);
        tfull = sl_lds32(sp_s + F::OFF_TM) == F::TFULL;
        unsigned a = sp_s;
#pragma unroll 1
        for (int half = 0; half < SL_COLS / 32; ++half) {
            {   // ---- B ring: prefetch rows [vb+LA, vb+LA+32), retire rows [vb-64, vb-32) ----
                const i64 vb = (i64)s * SL_COLS + half * 32;
                const i64 vl = vb + SL_LA + lane;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (vl < n) sl_cp8(bring[r] + (int)(vl & (SL_RB - 1)), bcol[r] + rowof(vl));
                sl_commit();
                sl_wait<SL_LA / 32 - 2>();
                __syncwarp();
                const i64 vs = vb - 64 + lane;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (vb >= 64 && vs < n && live[r]) bcol[r][rowof(vs)] = bring[r][(int)(vs & (SL_RB - 1))];
            }
#pragma unroll 1
            for (int g8 = 0; g8 < 32 / P; ++g8) {
                {   // pivot exchange
                    const unsigned pi8 = __byte_perm(code, 0u, 0x4441);  // byte 1 of code: 8 * (step + 1), 0 = not a pivot lane
                    __syncwarp();
#pragma unroll
                    for (int r = 0; r < R; ++r) sl_sts64(pbuf_s[r] + 8u + pi8, w[r]);
                    __syncwarp();
                }
                double v[R][P];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int c = 0; c < P; c += 2) {
                        const double2 p2 = sl_lds128(pbuf_s[r] + 16u + 8u * c);
                        v[r][c] = p2.x, v[r][c + 1] = p2.y;
                    }
                // ---- the triangle's multipliers, then the P steps; everything the chain does not need is issued behind them ----
                SlotRegs<P, BWD> sr;
                sl_fetch<P, BWD>(sr, a);
                double xs[R][P];
                if (tfull) {
#pragma unroll
                    for (int r = 0; r < R; ++r) sl_triangle<P, BWD, false>(v[r], xs[r], sr);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) sl_triangle<P, BWD, true>(v[r], xs[r], sr);
                }
                double2 m[P / 2];
#pragma unroll
                for (int c2 = 0; c2 < P / 2; ++c2) m[c2] = sl_lds128(a + F::OFF_M + 16u * (c2 * 32 + lane));
                // the row that enters through this lane (a pivot lane of this block: its old value went into pbuf above)
                if ((int)code < 0) {
#pragma unroll
                    for (int r = 0; r < R; ++r) w[r] = sl_lds64(bring_s[r] + (code >> 16 & 0xfffu));
                }
                const unsigned code_next = sl_lds32(a + F::BLK + F::OFF_CODE + 4u * lane);  // (past the stage for its last block: replaced)
                const unsigned tm_next = sl_lds32(a + F::BLK + F::OFF_TM);
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int c = 0; c < P; c += 2) sl_sts128(xprev_s[r] + 8u * c, xsp[r][c], xsp[r][c + 1]);  // every lane holds the same values
#pragma unroll
                for (int c2 = 0; c2 < P / 2; ++c2)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        sl_fma_if(w[r], xs[r][2 * c2], m[c2].x, code & (1u << (2 * c2)));
                        sl_fma_if(w[r], xs[r][2 * c2 + 1], m[c2].y, code & (2u << (2 * c2)));
                    }
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int c = 0; c < P; ++c) xsp[r][c] = xs[r][c];
                    xprev_s[r] = bring_s[r] + vbr8;
                }
                code = code_next;
                tfull = tm_next == F::TFULL;
                a += F::BLK;
                vbr8 = (vbr8 + 8u * P) & (8u * SL_RB - 1);
            }
        }
        // ---- end of a schedule stage: renormalise the slots (forward), hand the stage back ----
        if (!BWD) {
            const int em = (int)(sl_lds32(sp_s + F::OFF_END + (lane & ~3u)) >> (8 * (lane & 3))) & 31;
#pragma unroll
            for (int r = 0; r < R; ++r) w[r] = __shfl_sync(SL_FULL, w[r], em);
        }
        __syncwarp();
        if (lane == 0) sl_mbar_arrive(&empty[slot]);
        if (++slot == SL_NS) { slot = 0; phase ^= 1u; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < P; c += 2) sl_sts128(xprev_s[r] + 8u * c, xsp[r][c], xsp[r][c + 1]);  // the last block's solution entries
    // ---- flush the finished rows the in-loop stores have not reached ----
    __syncwarp();
    {
        i64 v0 = (i64)nstages * SL_COLS - 128;
        if (v0 < 0) v0 = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (live[r])
                for (i64 v = v0 + lane; v < n; v += 32) bcol[r][rowof(v)] = bring[r][(int)(v & (SL_RB - 1))];
    }
}

template <int P>
static size_t slot_sched_bytes(i64 n)
{
    const i64 nstages = cdiv64(n, SL_COLS);
    return (size_t)nstages * SlotFmt<P>::STAGE;
}

template <int P, bool BWD, int W, int R>
static int launch_slot_sweep(bmb200_ctx *h, i64 n, int kl, i64 nrhs, const unsigned char *sched, double *dB, i64 ldb)
{
    using F = SlotFmt<P>;
    const i64 nstages = cdiv64(n, SL_COLS);
    if (nstages >= ((i64)1 << 31)) return 1;
    const size_t smem = (size_t)SL_NS * F::STAGE + 2 * SL_NS * sizeof(unsigned long long) +
                        (size_t)W * R * (SL_RB * sizeof(double) + 8 * P + 16) + 64;
    BMB_CUDA(h, cudaFuncSetAttribute(gbtrs_slot<P, BWD, W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)cdiv64(nrhs, W * R);
    gbtrs_slot<P, BWD, W, R><<<blocks, (W + 1) * 32, smem, h->stream>>>(n, kl, nrhs, sched, (int)nstages, dB, ldb);
    BMB_LAUNCH_CHECK(h);
    return 0;
}

// PF / PB: steps per round of the forward / backward sweep; W warps per CTA; RF / RB right-hand sides per warp
template <int PF, int PB, int W, int RF, int RB>
static int run_slot(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv, double *dB,
                    i64 ldb)
{
    const i64 kv = kl + ku;
    const i64 nstages = cdiv64(n, SL_COLS);
    const size_t bf = slot_sched_bytes<PF>(n), bb = slot_sched_bytes<PB>(n);
    if (int rc = bmb_ensure_scratch(h, bf > bb ? bf : bb)) return rc;
    unsigned char *sched = static_cast<unsigned char *>(h->scratch);  // the two sweeps reuse the buffer (stream order)
    if (kl > 0) {
        slot_build_fwd<PF><<<(unsigned)cdiv64(nstages, 4), 128, 0, h->stream>>>(n, (int)kl, (int)kv, dAB, ldab, d_ipiv, sched, nstages);
        BMB_LAUNCH_CHECK(h);
        if (int rc = launch_slot_sweep<PF, false, W, RF>(h, n, (int)kl, nrhs, sched, dB, ldb)) return rc;
    }
    const i64 nblocks = nstages * SlotFmt<PB>::G;
    const i64 grid = imin64(cdiv64(nblocks, 4), (i64)h->sm_count * 16);
    slot_build_bwd<PB><<<(unsigned)grid, 128, 0, h->stream>>>(n, (int)kv, dAB, ldab, sched, nblocks);
    BMB_LAUNCH_CHECK(h);
    return launch_slot_sweep<PB, true, W, RB>(h, n, (int)kl, nrhs, sched, dB, ldb);
}

template <int PF, int PB, int RF, int RB>
static int run_slot_w(bmb200_ctx *h, int W, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv,
                      double *dB, i64 ldb)
{
    switch (W) {
    case 1: return run_slot<PF, PB, 1, RF, RB>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    case 2: return run_slot<PF, PB, 2, RF, RB>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    default: return run_slot<PF, PB, 4, RF, RB>(h, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    }
}

// Shapes this file covers with PF steps per forward round: the forward window (kl carried + PF entering rows) and the
// backward window (kl+ku rows, one per lane) both have to fit the 32 lanes of a warp.
static bool slot_covers(int PF, i64 n, i64 kl, i64 ku) { return n >= 1 && kl + PF <= 32 && kl + ku <= 32; }

// Returns 1 when the shape is not covered (the caller falls through to the older kernels).
// Measured at C4 (n = 2^20, (16,16), 256 RHS; tools/time_gbtrs.py): PF=4/PB=8, one right-hand side per warp, one warp
// per CTA: 92 ms; two / four right-hand sides per warp (interleaved chains in one instruction stream) cost 1.5x / 2.7x
// per warp -- no gain while the right-hand sides do not fill the chip's 592 sub-partition slots, and beyond that the
// hardware interleaves resident warps by itself -- so R stays 1.
int bmb_gbtrs_slot(bmb200_ctx *h, i64 n, i64 kl, i64 ku, i64 nrhs, const double *dAB, i64 ldab, const i64 *d_ipiv, double *dB,
                   i64 ldb)
{
    if (!slot_covers(4, n, kl, ku)) return 1;
    const i64 sms = h->sm_count;
    const int W = (nrhs <= 4 * sms) ? 1 : (nrhs <= 8 * sms) ? 2 : 4;
    return run_slot_w<4, 8, 1, 1>(h, W, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
}

// ---- development hook (include/bmb200_internal.h): explicit variant for A/B timing; not part of the public ABI ----
extern "C" int bmb200_internal_gbtrs_slot(bmb200_handle_t h, int PF, int PB, int W, int RF, int RB, int64_t n, int64_t kl,
                                          int64_t ku, int64_t nrhs, const double *dAB, int64_t ldab, const int64_t *d_ipiv,
                                          double *dB, int64_t ldb)
{
    if (!h) return -1;
    if (!(W == 1 || W == 2 || W == 4)) return -4;
    if (!slot_covers(PF, n, kl, ku) || nrhs < 1) return -7;
    DeviceGuard g(h->device);
#define SLOT_VARIANT(pf, pb, rf, rb) \
    if (PF == pf && PB == pb && RF == rf && RB == rb) return run_slot_w<pf, pb, rf, rb>(h, W, n, kl, ku, nrhs, dAB, ldab, d_ipiv, dB, ldb);
    SLOT_VARIANT(2, 2, 1, 1)
    SLOT_VARIANT(4, 4, 1, 1)
    SLOT_VARIANT(4, 8, 1, 1)
    SLOT_VARIANT(8, 8, 1, 1)
    SLOT_VARIANT(4, 4, 2, 2)
#undef SLOT_VARIANT
    return -2;
}

// ---- test hook: the verified two-term reciprocal division exactly as the backward sweep forms it ----
__global__ void sl_divcheck_kernel(i64 n, const double *__restrict__ x, const double *__restrict__ d, unsigned long long *__restrict__ bad)
{
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < n; k += (i64)gridDim.x * blockDim.x) {
        const double dv = d[k], t = x[k];
        double rh, rl;
        if (gb_div_safe_divisor(dv)) {
            rh = 1.0 / dv;
            rl = __dmul_rn(fma(-dv, rh, 1.0), rh);
        } else {
            rh = rl = __longlong_as_double(0x7ff8000000000000ll);
        }
        const double q = fma(t, rh, __dmul_rn(t, rl));
        const double q2 = fma(fma(-q, dv, t), rh, q);
        const bool ok = (q2 == q) && gb_exp_mid(t);
        const double got = ok ? q : ((t == 0.0 && rh == rh) ? __dmul_rn(t, rh) : gb_div_ieee(t, dv));
        const double ref = t / dv;
        if (__double_as_longlong(got) != __double_as_longlong(ref) && !(got != got && ref != ref)) atomicAdd(bad, 1ull);
        if (ok) atomicAdd(bad + 1, 1ull);
    }
}
extern "C" int bmb200_internal_divcheck2(bmb200_handle_t h, int64_t n, const double *dx, const double *dd, unsigned long long *dbad)
{
    if (!h) return -1;
    DeviceGuard g(h->device);
    sl_divcheck_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(n, dx, dd, dbad);
    BMB_LAUNCH_CHECK(h);
    return 0;
}
