"""Lane-level numpy emulation of the DMMA (mma.sync m8n8k4 f64) Riccati stage factorization used by
ipm_srbd.cuh (sweep_factor).  Every "register" is a length-32 array (one value per lane); dmma() applies the PTX
fragment layout of mma.m8n8k4.f64 (A: row = lane>>2, col = lane&3; B: row = lane&3, col = lane>>2;
C/D: row = lane>>2, cols 2*(lane&3)+{0,1}).  Validates the fragment / permutation bookkeeping on the CPU
against a plain numpy Riccati step before it is transcribed to CUDA.

    python scripts/proto_dmma_factor.py        (also run by tests/test_fragment_protos.py)
"""
import numpy as np

LANES = np.arange(32)
R_, T_ = LANES >> 2, LANES & 3
PI = (R_ >> 1) + 4 * (R_ & 1)  # row permutation of the B operand: c0 -> columns 0..3, c1 -> columns 4..7


def dmma(c0, c1, a, b):
    A = np.zeros((8, 4)); Bm = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[R_, T_] = a
    Bm[T_, R_] = b
    C[R_, 2 * T_] = c0
    C[R_, 2 * T_ + 1] = c1
    D = A @ Bm + C
    return D[R_, 2 * T_].copy(), D[R_, 2 * T_ + 1].copy()


def frag(X, I, kt, nrows=None):
    """lane(r,t) holds X[8I+r][4kt+t] (0 outside)."""
    out = np.zeros(32)
    rows = 8 * I + R_; cols = 4 * kt + T_
    ok = (rows < X.shape[0]) & (cols < X.shape[1])
    if nrows is not None:
        ok &= rows < nrows
    out[ok] = X[rows[ok], cols[ok]]
    return out


def pfrag(X, J, kt, nrows=None):
    out = np.zeros(32)
    rows = 8 * J + PI; cols = 4 * kt + T_
    ok = (rows < X.shape[0]) & (cols < X.shape[1])
    if nrows is not None:
        ok &= rows < nrows
    out[ok] = X[rows[ok], cols[ok]]
    return out


def factor_stage(G, Pn, Hinit, gt, n, reg):
    """G: (24,12) rows [B^T; A^T]; Pn: P_{k+1} (12,12) symmetric; Hinit: (24,24) lower = H + D^T Gamma D;
    gt: gradient (24) (already rg + D^T gamma + G (P rb + p)); n = 24 or 12 (stage 0).
    Returns L (n x 12 lower-trapezoid), Linv (12x12), P (12x12 or None), lv (12), p (12 or None)."""
    nI = 3 if n == 24 else 2
    # ---- fragments ----
    GF = [[frag(G, I, kt, n) for kt in range(3)] for I in range(nI)]
    GPF = [[pfrag(G, J, kt, n) for kt in range(3)] for J in range(nI)]
    PPF = [[pfrag(Pn, Jn, kt, 12) for kt in range(3)] for Jn in range(2)]
    # ---- AL = G P ----
    ALF = [[None] * 3 for _ in range(nI)]
    for I in range(nI):
        c0 = np.zeros(32); c1 = np.zeros(32)
        for kt in range(3):
            c0, c1 = dmma(c0, c1, GF[I][kt], PPF[0][kt])
        ALF[I][0], ALF[I][1] = c0, c1
        c0 = np.zeros(32); c1 = np.zeros(32)
        for kt in range(3):
            c0, c1 = dmma(c0, c1, GF[I][kt], PPF[1][kt])
        ALF[I][2] = c0  # c1 = columns 12..15 of AL: padding
    # ---- M = Hinit + AL G^T (lower tiles), MF[I][p] ----
    MF = [[None] * 6 for _ in range(3)]
    for I in range(nI):
        for J in range(I + 1):
            c0 = np.zeros(32); c1 = np.zeros(32)
            for kt in range(3):
                c0, c1 = dmma(c0, c1, ALF[I][kt], GPF[J][kt])
            MF[I][2 * J] = c0 + frag(Hinit, I, 2 * J)
            MF[I][2 * J + 1] = c1 + frag(Hinit, I, 2 * J + 1)
    # reg on the diagonal (row == col lanes)
    for I in range(nI):
        for p in (2 * I, 2 * I + 1):
            diag = (8 * I + R_) == (4 * p + T_)
            MF[I][p] = MF[I][p] + np.where(diag & (8 * I + R_ < n), reg, 0.0)
    # ---- E rows: per panel p the identity block of rows 4p..4p+3 (the substitution turns it into L_pp^-T) ----
    EF = [[np.where((8 * e + R_ == 4 * p + T_) & (8 * e + R_ < 12), 1.0, 0.0) for p in range(3)] for e in range(2)]
    FT = np.zeros((3, 25, 4))
    # panel buffer rows: 0..23 M rows, 24 gradient, 25..36 E rows
    g = gt.copy()  # lane c holds g[c] (row-per-lane), here a plain vector
    Lrows = np.zeros((37, 12))
    inv_diag = np.zeros(12)
    for p in range(3):
        pan = np.zeros((37, 4))
        # STS frags of panel p
        for I in range(p >> 1, nI):
            pan[8 * I + R_, T_] = MF[I][p]
        for e in range(2):
            rows = 8 * e + R_
            ok = rows < 12
            pan[25 + rows[ok], T_[ok]] = EF[e][p][ok]
        pan[24, :] = g[4 * p:4 * p + 4]
        # ---- every lane: 4x4 Cholesky of the diagonal block (lower), redundantly ----
        a = pan[4 * p:4 * p + 4, :]
        L44 = np.zeros((4, 4)); iv = np.zeros(4)
        for j in range(4):
            dj = a[j, j]
            for l in range(j):
                dj -= L44[j, l] * L44[j, l]
            iv[j] = 1.0 / np.sqrt(dj)
            L44[j, j] = dj * iv[j]
            for i in range(j + 1, 4):
                s = a[i, j]
                for l in range(j):
                    s -= L44[i, l] * L44[j, l]
                L44[i, j] = s * iv[j]
        inv_diag[4 * p:4 * p + 4] = iv
        # ---- own-row substitution: rows 4p..23 (M), 24 (gradient), E rows 0..4p+3 ----
        rows = [i for i in range(4 * p, n)] + [24] + [25 + i for i in range(4 * p, 4 * p + 4)]
        for row in rows:
            a0 = pan[row]
            l = np.zeros(4)
            for j in range(4):
                s = a0[j]
                for q in range(j):
                    s -= l[q] * L44[j, q]
                l[j] = s * iv[j]
            pan[row] = l
            Lrows[row, 4 * p:4 * p + 4] = l
        # diagonal of the block: dj * inv (the substitution gives the same value)
        # ---- gradient row update (vector): g[c] -= sum_l lv[l] L[c][l], c > 4p+3 ----
        lvp = pan[24]
        for c in range(4 * p + 4, n):
            g[c] -= float(np.dot(lvp, pan[c]))  # order irrelevant for the check
        # ---- trailing updates with DMMA: A = -L panel frag, B = permuted frag ----
        def Afrag(rowbase):
            out = np.zeros(32)
            rows_ = rowbase + R_
            ok = rows_ < 37
            out[ok] = -pan[rows_[ok], T_[ok]]
            return out

        def Bfrag(J):
            return pan[8 * J + PI, T_].copy()
        Jlist = {0: [0, 1, 2], 1: [1, 2], 2: [1, 2]}[p]
        for I in range(nI):
            for J in Jlist:
                if J > I or J >= nI:
                    continue
                if 8 * I + 7 < 4 * p + 4:
                    continue
                d0, d1 = dmma(MF[I][2 * J], MF[I][2 * J + 1], Afrag(8 * I), Bfrag(J))
                if 2 * J > p:
                    MF[I][2 * J] = d0
                if 2 * J + 1 > p:
                    MF[I][2 * J + 1] = d1
        # factor panel p as the kernel exports it: rows 4p..4p+3 = L_pp^-T (E rows), rows > 4p+3 (< 12) = L rows,
        # rows 12..23 = Ls, row 24 = lv
        for i in range(12):
            if 4 * p <= i < 4 * p + 4:
                FT[p, i] = pan[25 + i]
            elif i >= 4 * p + 4:
                FT[p, i] = pan[i]
        FT[p, 12:24] = pan[12:24] if n == 24 else 0.0
        FT[p, 24] = pan[24]
    L = Lrows[:n, :]
    lv = Lrows[24, :]
    # what the vector sweeps do with the panels: blocked forward substitution = L^-1 (checked column by column)
    Linv = np.zeros((12, 12))
    for c in range(12):
        e_c = np.zeros(12); e_c[c] = 1.0
        out = np.zeros(12)
        for pb in range(3):
            gp = e_c[4 * pb:4 * pb + 4].copy()
            for qb in range(pb):
                gp -= FT[qb, 4 * pb:4 * pb + 4, :] @ out[4 * qb:4 * qb + 4]      # L_pq block
            out[4 * pb:4 * pb + 4] = FT[pb, 4 * pb:4 * pb + 4, :].T @ gp        # (L_pp^-T)^T
        Linv[:, c] = out
    LinvT = Linv.T
    if n == 24:
        P = np.zeros((12, 12))
        for I, p in ((1, 3), (2, 3), (2, 4), (2, 5)):
            rows_ = 8 * I + R_; cols_ = 4 * p + T_
            ok = (rows_ >= 12) & (cols_ <= rows_)
            P[rows_[ok] - 12, cols_[ok] - 12] = MF[I][p][ok]
            P[cols_[ok] - 12, rows_[ok] - 12] = MF[I][p][ok]
        return L, LinvT.T, P, lv, g[12:24]
    return L, LinvT.T, None, lv, None


def reference(G, Pn, Hinit, gt, n, reg):
    Gn = G[:n]
    H = np.tril(Hinit[:n, :n]); H = H + np.tril(H, -1).T
    M = H + Gn @ Pn @ Gn.T + reg * np.eye(n)
    Lr = np.linalg.cholesky(M[:12, :12])
    Linv = np.linalg.inv(Lr)
    lv = Linv @ gt[:12]
    if n == 24:
        Ls = M[12:, :12] @ Linv.T
        P = M[12:, 12:] - Ls @ Ls.T
        p = gt[12:24] - Ls @ lv
        return np.vstack([Lr, Ls]), Linv, P, lv, p
    return Lr, Linv, None, lv, None


def main():
    rng = np.random.default_rng(0)
    for n in (24, 12):
        G = rng.normal(size=(24, 12))
        X = rng.normal(size=(12, 12)); Pn = X @ X.T + np.eye(12)
        Hinit = np.zeros((24, 24))
        Y = rng.normal(size=(12, 12)); Hinit[:12, :12] = np.tril(Y @ Y.T + np.eye(12))
        Hinit[:12, :12] += np.triu(np.full((12, 12), np.nan), 1)  # upper garbage must not matter
        Hinit[12:, 12:] = np.diag(rng.uniform(1, 2, 12))
        gt = rng.normal(size=24)
        Hclean = np.nan_to_num(Hinit)
        # the kernel loads only the lower triangle: emulate by zeroing the upper garbage
        out = factor_stage(G, Pn, np.tril(Hclean), gt, n, 1e-12)
        ref = reference(G, Pn, Hclean, gt, n, 1e-12)
        for name, a, b in zip(("L", "Linv", "P", "lv", "p"), out, ref):
            if a is None:
                continue
            if name == "L":
                a = a.copy(); a[:12] = np.tril(a[:12])
            err = np.max(np.abs(a - b)) / np.max(np.abs(b))
            print("n=%d %-5s rel err %.2e" % (n, name, err))
            assert err < 1e-11, name


if __name__ == "__main__":
    main()
