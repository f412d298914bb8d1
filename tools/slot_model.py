"""CPU model of the slot-scheduled band solve (csrc/gbtrs_slot.cu): builds the per-block schedule exactly as the device
prepass does and replays it with scalar arithmetic, then compares with the oracle's DGBTRS.  Development aid only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from oracle import brand, ldiv, lu  # noqa: E402

NL = 32


def fma(a, b, c):
    from fractions import Fraction

    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def build_fwd(n, kl, kv, ab, ipiv, P, G):
    """One record per block: src[P], ent[32] (row offset loaded AFTER the block's shuffles into the lane, 0xFF none),
    om[32], M[32][P], T[i][c], tm; per stage: endmap[32].  Rows j0 .. j0+kl+P-1 sit in lanes 0 .. kl+P-1 at a stage start;
    the rows entering the window during block g+1 are loaded at the end of block g into the lanes block g's pivots freed."""
    nblk = -(-n // P)
    nst = -(-nblk // G)
    blocks = []
    endmaps = []
    for s in range(nst):
        j0 = s * G * P
        pos = [-1] * NL
        for i in range(kl + P):
            pos[i] = j0 + i if j0 + i < n else -1
        for g in range(G):
            jb = j0 + g * P
            src = [0] * P
            real = [False] * P
            ent = [0xFF] * NL
            om = [0] * NL
            M = np.zeros((NL, P))
            for c in range(P):
                j = jb + c
                if j >= n:
                    continue
                real[c] = True
                piv = int(ipiv[j]) - 1
                s_ = pos.index(piv)
                f_ = pos.index(j)
                src[c] = s_
                if s_ != f_:
                    pos[f_] = piv
                pos[s_] = -1
                km = min(kl, n - 1 - j)
                for q in range(NL):
                    if pos[q] > j and pos[q] <= j + km:
                        o = pos[q] - j - 1
                        M[q, c] = ab[kv + 1 + o, j]
                        om[q] |= 1 << c
            T = np.zeros((P, P))
            tm = np.zeros((P, P), dtype=bool)
            for i in range(P):
                for c in range(i):
                    if real[i]:
                        T[i, c] = M[src[i], c]
                        tm[i, c] = bool(om[src[i]] >> c & 1)
            for c in range(P):
                if real[c]:
                    om[src[c]] = 0          # a pivot lane's register is dead after the shuffle: it takes an entering row
                    row = jb + P + kl + c
                    if row < n:
                        pos[src[c]] = row
                        ent[src[c]] = P + kl + c
            blocks.append(dict(jb=jb, src=src, ent=ent, om=om, M=M, T=T, tm=tm))
        jend = j0 + G * P
        em = list(range(NL))
        for i in range(kl + P):
            if jend + i < n:
                em[i] = pos.index(jend + i)
        endmaps.append(em)
    return blocks, endmaps


def solve_fwd(n, kl, b, blocks, endmaps, P, G):
    w = [0.0] * NL
    for i in range(min(kl + P, n)):
        w[i] = b[i]
    x = b.copy()
    for bi, B in enumerate(blocks):
        jb = B["jb"]
        v = [w[B["src"][c]] for c in range(P)]
        for q in range(NL):
            if B["ent"][q] != 0xFF:
                w[q] = b[jb + B["ent"][q]]
        xs = [0.0] * P
        for i in range(P):
            t = v[i]
            for c in range(i):
                if B["tm"][i, c]:
                    t = fma(-xs[c], B["T"][i, c], t)
            xs[i] = t
        for q in range(NL):
            for c in range(P):
                if B["om"][q] >> c & 1:
                    w[q] = fma(-xs[c], B["M"][q, c], w[q])
        for c in range(P):
            if jb + c < n:
                x[jb + c] = xs[c]
        if (bi + 1) % G == 0:
            em = endmaps[bi // G]
            w = [w[em[q]] for q in range(NL)]
    return x


def build_bwd(n, kv, ab, P):
    """virtual index v = n-1-row; lane = v % 32."""
    nblk = -(-n // P)
    blocks = []
    for bI in range(nblk):
        vb = bI * P
        src = [(vb + c) % NL for c in range(P)]
        ent = [0xFF] * NL
        om = [0] * NL
        M = np.zeros((NL, P))
        d = np.ones(P)
        rowv = [0] * NL
        for q in range(NL):
            # the lane's row (virtual) after the post-shuffle entry: unique v == q mod 32 in [vb+P, vb+P+32)
            v = vb + P + ((q - (vb + P)) % NL)
            rowv[q] = v
        for c in range(P):
            ent[(vb + c) % NL] = NL + c if vb + c + NL < n else 0xFF
        for c in range(P):
            vp = vb + c
            if vp >= n:
                continue
            col = n - 1 - vp
            d[c] = ab[kv, col]
            for q in range(NL):
                dist = rowv[q] - vp
                if 1 <= dist <= kv and rowv[q] < n:
                    M[q, c] = ab[kv - dist, col]
                    om[q] |= 1 << c
        T = np.zeros((P, P))
        tm = np.zeros((P, P), dtype=bool)
        for i in range(P):
            for c in range(i):
                if vb + i < n and (i - c) <= kv:
                    col = n - 1 - (vb + c)
                    T[i, c] = ab[kv - (i - c), col]
                    tm[i, c] = True
        blocks.append(dict(vb=vb, src=src, ent=ent, om=om, M=M, T=T, tm=tm, d=d))
    return blocks


def solve_bwd(n, b, blocks, P):
    bv = b[::-1].copy()  # virtual order
    x = bv.copy()
    w = [bv[q] if q < n else 0.0 for q in range(NL)]
    for B in blocks:
        vb = B["vb"]
        v = [w[B["src"][c]] for c in range(P)]
        for q in range(NL):
            if B["ent"][q] != 0xFF:
                w[q] = bv[vb + B["ent"][q]]
        xs = [0.0] * P
        for i in range(P):
            t = v[i]
            for c in range(i):
                if B["tm"][i, c]:
                    t = fma(-xs[c], B["T"][i, c], t)
            xs[i] = t / B["d"][i]
        for q in range(NL):
            for c in range(P):
                if B["om"][q] >> c & 1:
                    w[q] = fma(-xs[c], B["M"][q, c], w[q])
        for c in range(P):
            if vb + c < n:
                x[vb + c] = xs[c]
    return x[::-1].copy()


def main():
    rng = np.random.default_rng(5)
    be = oracle.backend("C")
    for (n, l, u, P, G) in [(200, 16, 16, 4, 16), (77, 4, 3, 4, 4), (130, 5, 7, 8, 2), (64, 3, 2, 4, 16), (300, 24, 8, 8, 8),
                            (50, 1, 1, 4, 16), (5, 4, 4, 4, 16), (1, 0, 0, 4, 16), (260, 28, 4, 4, 16), (100, 0, 3, 4, 16),
                            (100, 3, 0, 4, 16)]:
        A = brand(rng, n, n, l, u)
        ab, ipiv, info = lu(be, A)
        kv = l + u  # LAPACK ku of the factor storage is u; kv = kl+ku
        b = rng.standard_normal(n)
        ref = np.asfortranarray(b.reshape(n, 1).copy())
        ldiv(be, "N", ab, ipiv, l, u, ref)
        y = b.copy()
        if l > 0:
            blocks, endmaps = build_fwd(n, l, kv, ab, ipiv, P, G)
            y = solve_fwd(n, l, b, blocks, endmaps, P, G)
        bb = build_bwd(n, kv, ab, P)
        x = solve_bwd(n, y, bb, P)
        ok = np.array_equal(x, ref[:, 0])
        print((n, l, u, P, G), "bit-identical" if ok else f"MISMATCH max {np.max(np.abs(x - ref[:, 0]))}")
        assert ok


if __name__ == "__main__":
    main()
