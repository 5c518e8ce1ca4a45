"""Lane-level numpy emulation of the DMMA "vector-fragment" form of the vector sweeps of ipm_srbd.cuh
(S4 backward gradient recursion, S2 forward rollout): a length-12 vector is 3 registers valid in lanes 0..3
(lane t of k-tile kt holds v[4 kt + t]) = row 0 of the A operand; y = A x is 2 x 3 mma.m8n8k4 with the matrix as
row-permuted B fragments, and the accumulator (c0, c1) of output tile I IS the next A operand (k-tiles 2I, 2I+1).
Checks the fragment gather addresses (factor panels, P, BAbt panel-major) against plain numpy.

    python scripts/proto_dmma_vec.py        (also run by tests/test_fragment_protos.py)
"""
import numpy as np
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from proto_dmma_factor import dmma, R_, T_, PI  # noqa: E402

KPANF = 102


def vfrag(v, kt):
    """A operand holding the vector in row 0: every lane loads v[4kt + t] (rows r > 0 replicate; irrelevant)."""
    idx = 4 * kt + T_
    return np.where(idx < len(v), v[np.minimum(idx, len(v) - 1)], 0.0)


def vout(c0, c1, I, n):
    """accumulator of output tile I -> entries 8I+t (c0), 8I+4+t (c1) read from lanes r == 0"""
    out = {}
    for t in range(4):
        if 8 * I + t < n:
            out[8 * I + t] = c0[t]
        if 8 * I + 4 + t < n:
            out[8 * I + 4 + t] = c1[t]
    return out


def gather(mem, addr, ok):
    a = np.where(ok, addr, 0)
    return np.where(ok, mem[a], 0.0)


def main():
    rng = np.random.default_rng(1)
    Linv = np.tril(rng.normal(size=(12, 12)))
    Ls = rng.normal(size=(12, 12))
    lv = rng.normal(size=12)
    X = rng.normal(size=(12, 12)); P = X @ X.T
    G = rng.normal(size=(24, 12))
    # ---- storage: factor panels FT[3][25][4], P row-major, BAbt panel-major (28 x 12) ----
    # factor panels as the factorization exports them: panel p, rows 0..11: rows 4p..4p+3 = L_pp^-T, rows > 4p+3 = L rows;
    # rows 12..23 = Ls, row 24 = lv  (L = inverse of the random Linv)
    Lm = np.linalg.inv(Linv)
    FT = np.zeros(3 * KPANF)
    for pb in range(3):
        Xp = np.linalg.inv(Lm[4 * pb:4 * pb + 4, 4 * pb:4 * pb + 4])
        for c in range(4):
            for i in range(4 * pb, 4 * pb + 4):
                FT[pb * KPANF + i * 4 + c] = Xp.T[i - 4 * pb][c]
            for i in range(4 * pb + 4, 12):
                FT[pb * KPANF + i * 4 + c] = Lm[i][4 * pb + c]
            for i in range(12):
                FT[pb * KPANF + (12 + i) * 4 + c] = Ls[i][4 * pb + c]
            FT[pb * KPANF + 96 + c] = lv[4 * pb + c]
    Pm = P.reshape(-1).copy()
    BAbt = np.zeros(336)
    for i in range(24):
        for l in range(12):
            BAbt[(i >> 2) * 48 + 4 * l + (i & 3)] = G[i][l]
    r, t, pi = R_, T_, PI

    def frags(addr_fn, mem, nI, nK, okfn):
        return [[gather(mem, addr_fn(I, kt), okfn(I, kt)) for kt in range(nK)] for I in range(nI)]
    # output row index i = 8I + pi(r), input index j = 4kt + t
    i_of = lambda I: 8 * I + pi
    ok12 = lambda I, kt: i_of(I) < 12
    F_Ls = frags(lambda I, kt: kt * KPANF + (12 + i_of(I)) * 4 + t, FT, 2, 3, ok12)
    F_LsT = frags(lambda I, kt: (i_of(I) >> 2) * KPANF + (12 + 4 * kt + t) * 4 + (i_of(I) & 3), FT, 2, 3, ok12)
    # 4x4 blocks of the panels (the kernel's bA / bB offsets): row index t / column pi, and row pi / column t
    bA = 4 * t + (pi & 3); bB = 4 * (pi & 3) + t; ok4 = pi < 4
    blkA = lambda col_panel, row_block: gather(FT, bA + col_panel * KPANF + 16 * row_block, ok4)
    blkB = lambda col_panel, row_block: gather(FT, bB + col_panel * KPANF + 16 * row_block, ok4)
    F_P = frags(lambda I, kt: i_of(I) * 12 + 4 * kt + t, Pm, 2, 3, ok12)
    F_G = frags(lambda I, kt: (i_of(I) >> 2) * 48 + 4 * (4 * kt + t) + (i_of(I) & 3), BAbt, 3, 3, lambda I, kt: i_of(I) < 24)
    F_GT = frags(lambda I, kt: kt * 48 + 4 * i_of(I) + t, BAbt, 2, 6, ok12)   # G^T: out j = 8I+pi, in i = 4kt+t

    def gemv(F, xk, cinit, nI):
        """y tile I = cinit[I] + sum_kt dmma(x k-tile kt, F[I][kt]); xk: list of A operands"""
        out = []
        for I in range(nI):
            c0, c1 = cinit[I]
            for kt in range(len(xk)):
                c0, c1 = dmma(c0, c1, xk[kt], F[I][kt])
            out.append((c0, c1))
        return out

    def to_vec(tiles, n):
        v = np.zeros(n)
        for I, (c0, c1) in enumerate(tiles):
            for k_, val in vout(c0, c1, I, n).items():
                v[k_] = val
        return v

    def cfrag(v, I):
        """C init of output tile I from a plain vector: lanes load v[8I + t], v[8I + 4 + t]"""
        i0 = 8 * I + T_; i1 = 8 * I + 4 + T_
        return (np.where(i0 < len(v), v[np.minimum(i0, len(v) - 1)], 0.0),
                np.where(i1 < len(v), v[np.minimum(i1, len(v) - 1)], 0.0))
    Z = np.zeros(32)
    # ===== S4 body =====
    rg = rng.normal(size=24); tvec = rng.normal(size=12); dtg = rng.normal(size=12)  # dtg = D^T gamma (u rows)
    gt = [cfrag(rg, I) for I in range(3)]
    gt = gemv(F_G, [vfrag(tvec, kt) for kt in range(3)], gt, 3)          # g~ = rg + G t
    g_ref = rg + G @ tvec
    assert np.allclose(to_vec(gt, 24), g_ref)
    # lv = L^-1 g_u: blocked forward substitution; one DMMA per 4x4 block, c0 of the accumulator = the k-tile
    gu = [gt[0][0], gt[0][1], gt[1][0]]
    lv0, _ = dmma(Z, Z, gu[0], blkA(0, 0))
    g1, _ = dmma(gu[1], Z, -lv0, blkB(0, 1))
    g2, _ = dmma(gu[2], Z, -lv0, blkB(0, 2))
    lv1, _ = dmma(Z, Z, g1, blkA(1, 1))
    g2, _ = dmma(g2, Z, -lv1, blkB(1, 2))
    lv2, _ = dmma(Z, Z, g2, blkA(2, 2))
    lvt = [(lv0, lv1), (lv2, Z)]
    lv_ref = Linv @ g_ref[:12]
    assert np.allclose(to_vec(lvt, 12), lv_ref)
    # p = g_x - Ls lv : C init = (tile1.c1, tile2.c0), (tile2.c1, -)
    nlv = [-lvt[0][0], -lvt[0][1], -lvt[1][0]]
    pt = gemv(F_Ls, nlv, [(gt[1][1], gt[2][0]), (gt[2][1], Z)], 2)
    p_ref = g_ref[12:] - Ls @ lv_ref
    assert np.allclose(to_vec(pt, 12), p_ref)
    # ===== S2 body =====
    x = rng.normal(size=12); rb = rng.normal(size=12); pn = rng.normal(size=12)
    xk = [vfrag(x, kt) for kt in range(3)]
    tt = gemv(F_LsT, xk, [cfrag(lv, 0), cfrag(lv, 1)], 2)                # t = Ls^T x + lv
    t_ref = Ls.T @ x + lv
    assert np.allclose(to_vec(tt, 12), t_ref)
    nt = [-tt[0][0], -tt[0][1], -tt[1][0]]
    # u = -L^-T t: blocked back substitution
    u2, _ = dmma(Z, Z, nt[2], blkB(2, 2))
    w1, _ = dmma(nt[1], Z, -u2, blkA(1, 2))
    w0, _ = dmma(nt[0], Z, -u2, blkA(0, 2))
    u1, _ = dmma(Z, Z, w1, blkB(1, 1))
    w0, _ = dmma(w0, Z, -u1, blkA(0, 1))
    u0, _ = dmma(Z, Z, w0, blkB(0, 0))
    ut = [(u0, u1), (u2, Z)]
    u_ref = -Linv.T @ t_ref
    assert np.allclose(to_vec(ut, 12), u_ref)
    zk = [ut[0][0], ut[0][1], ut[1][0]] + xk                             # z = [u; x] as 6 k-tiles
    xn = gemv(F_GT, zk, [cfrag(rb, 0), cfrag(rb, 1)], 2)                 # x+ = G^T z + rb
    xn_ref = G.T @ np.concatenate([u_ref, x]) + rb
    assert np.allclose(to_vec(xn, 12), xn_ref)
    xnk = [xn[0][0], xn[0][1], xn[1][0]]
    dpi = gemv(F_P, xnk, [cfrag(pn, 0), cfrag(pn, 1)], 2)                # dpi = P x+ + p
    assert np.allclose(to_vec(dpi, 12), P @ xn_ref + pn)
    print("S4 / S2 fragment-form bodies OK (blocked triangular solves)")


if __name__ == "__main__":
    main()
