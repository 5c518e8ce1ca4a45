"""Lane-level numpy emulation of the DMMA "vector-fragment" form of the vector sweeps of ipm_srbd.cuh
(S4 backward gradient recursion, S2 forward rollout): a length-12 vector is 3 registers valid in lanes 0..3
(lane t of k-tile kt holds v[4 kt + t]) = row 0 of the A operand; y = A x is 2 x 3 mma.m8n8k4 with the matrix as
row-permuted B fragments, and the accumulator (c0, c1) of output tile I IS the next A operand (k-tiles 2I, 2I+1).
Checks the fragment gather addresses (factor panels, P, BAbt panel-major) against plain numpy.

    python scripts/proto_dmma_vec.py
"""
import numpy as np
from proto_dmma_factor import dmma, R_, T_, PI

KPANF = 100


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
    FT = np.zeros(3 * KPANF)
    for l in range(12):
        for i in range(12):
            FT[(l >> 2) * KPANF + i * 4 + (l & 3)] = Linv[l][i]          # E row i, column l = Linv^T[i][l]
            FT[(l >> 2) * KPANF + (12 + i) * 4 + (l & 3)] = Ls[i][l]
        FT[(l >> 2) * KPANF + 96 + (l & 3)] = lv[l]
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
    F_Linv = frags(lambda I, kt: (i_of(I) >> 2) * KPANF + (4 * kt + t) * 4 + (i_of(I) & 3), FT, 2, 3, ok12)
    F_Ls = frags(lambda I, kt: kt * KPANF + (12 + i_of(I)) * 4 + t, FT, 2, 3, ok12)
    F_LsT = frags(lambda I, kt: (i_of(I) >> 2) * KPANF + (12 + 4 * kt + t) * 4 + (i_of(I) & 3), FT, 2, 3, ok12)
    F_LinvT = frags(lambda I, kt: kt * KPANF + i_of(I) * 4 + t, FT, 2, 3, ok12)
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
    # lv = Linv g_u : A operand k-tiles = (tile0.c0, tile0.c1, tile1.c0)
    gu = [gt[0][0], gt[0][1], gt[1][0]]
    lvt = gemv(F_Linv, gu, [(Z, Z), (Z, Z)], 2)
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
    ut = gemv(F_LinvT, nt, [(Z, Z), (Z, Z)], 2)                          # u = -Linv^T t
    u_ref = -Linv.T @ t_ref
    assert np.allclose(to_vec(ut, 12), u_ref)
    zk = [ut[0][0], ut[0][1], ut[1][0]] + xk                             # z = [u; x] as 6 k-tiles
    xn = gemv(F_GT, zk, [cfrag(rb, 0), cfrag(rb, 1)], 2)                 # x+ = G^T z + rb
    xn_ref = G.T @ np.concatenate([u_ref, x]) + rb
    assert np.allclose(to_vec(xn, 12), xn_ref)
    xnk = [xn[0][0], xn[0][1], xn[1][0]]
    dpi = gemv(F_P, xnk, [cfrag(pn, 0), cfrag(pn, 1)], 2)                # dpi = P x+ + p
    assert np.allclose(to_vec(dpi, 12), P @ xn_ref + pn)
    print("S4 / S2 fragment-form bodies OK")


if __name__ == "__main__":
    main()
