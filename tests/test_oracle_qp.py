"""CPU tests (-m "not gpu"): pin the oracle's OCP-QP solve against the reference's own golden vectors
and known-answer identities (hpipm-cpp/test/ocp_qp_ipm_solver.cpp)."""
import numpy as np
import pytest


def _dims(pkg, d):
    from srbd_nmpc_solver_b200.binding import make_dims
    return make_dims(**d)


def _unflat(a, r, c):
    """column-major flat [..., r*c] -> [..., r, c]"""
    return np.swapaxes(a.reshape(a.shape[:-1] + (c, r)), -1, -2)


def is_approx(a, b, prec):
    """Eigen's isApprox: ||a-b|| <= prec * min(||a||, ||b||)."""
    return np.linalg.norm(a - b) <= prec * min(np.linalg.norm(a), np.linalg.norm(b))


@pytest.mark.parametrize("shorten,expected_iters", [
    # SURVEY.md Appendix D, row "older shortening alpha*((1-alpha)*0.99+alpha*0.9999999)"
    (1, [17, 10, 8, 7, 6, 7, 5, 5, 4, 3, 4, 4, 4, 3, 3]),
    (0, None),
])
def test_compare_results_golden(pkg, orc, golden_quadcopter, shorten, expected_iters):
    """hpipm-cpp/test/ocp_qp_ipm_solver.cpp:170-315 (compareResults): 15 closed-loop MPC steps, each
    [x0..xN,u0..uN-1] must match sol{t}.txt with isApprox(1e-9) (:310)."""
    dims_d, arrays, settings, A, Bm = pkg.workload.quadcopter_mpc()
    dims = _dims(pkg, dims_d)
    args = orc.ipm_args(alpha_shorten=shorten, **settings)
    x = np.zeros(12)
    iters = []
    for t in range(15):
        arrays["x0"] = x[None, :].copy()
        out = orc.qp_solve(dims, args, arrays, 1)
        assert out["status"][0] == 0
        cat = np.concatenate([out["x"][0].reshape(-1), out["u"][0].reshape(-1)])
        assert is_approx(cat, golden_quadcopter[t], 1e-9), t
        assert np.allclose(out["x"][0, 0], x)
        iters.append(int(out["iter"][0]))
        arrays["x_init"], arrays["u_init"] = out["x"].copy(), out["u"].copy()  # warm_start = 1
        x = A @ x + Bm @ out["u"][0, 0]
    if expected_iters is not None:
        assert iters == expected_iters


@pytest.mark.parametrize("ric_alg", [0, 1])
def test_unconstrained_analytic_riccati(pkg, orc, ric_alg):
    """hpipm-cpp/test/ocp_qp_ipm_solver.cpp:22-110 (unconstrained): status Success, iter == 0 (:56),
    x,u,pi,P,-p,K,k equal a textbook Riccati recursion with isApprox(1e-10) (:92-109)."""
    B, N, nx, nu = 4, 20, 5, 3
    dims_d, arrays = pkg.workload.random_qp(B, N=N, nx=nx, nu=nu, seed=7)
    out = orc.qp_solve(_dims(pkg, dims_d), orc.ipm_args(ric_alg=ric_alg), arrays, B)
    assert (out["status"] == 0).all() and (out["iter"] == 0).all()
    for i in range(B):
        A = _unflat(arrays["A"][i], nx, nx); Bm = _unflat(arrays["Bm"][i], nx, nu)
        Q = _unflat(arrays["Q"][i], nx, nx); S = _unflat(arrays["S"][i], nu, nx)
        R = _unflat(arrays["R"][i], nu, nu)
        b, q, r, x0 = arrays["b"][i], arrays["q"][i], arrays["r"][i], arrays["x0"][i]
        P = [None] * (N + 1); s = [None] * (N + 1); K = [None] * N; kf = [None] * N
        P[N] = Q[N]; s[N] = -q[N]
        for k in range(N - 1, -1, -1):
            F = Q[k] + A[k].T @ P[k + 1] @ A[k]
            H = S[k] + Bm[k].T @ P[k + 1] @ A[k]
            G = R[k] + Bm[k].T @ P[k + 1] @ Bm[k]
            Gi = np.linalg.inv(G)
            K[k] = -Gi @ H
            kf[k] = -Gi @ (Bm[k].T @ P[k + 1] @ b[k] - Bm[k].T @ s[k + 1] + r[k])
            P[k] = F - K[k].T @ G @ K[k]
            s[k] = A[k].T @ (s[k + 1] - P[k + 1] @ b[k]) - q[k] - H.T @ kf[k]
        xs = [x0]; us = []
        for k in range(N):
            us.append(K[k] @ xs[k] + kf[k])
            xs.append(A[k] @ xs[k] + Bm[k] @ us[k] + b[k])
        prec = 1e-10
        for k in range(N + 1):
            assert is_approx(xs[k], out["x"][i, k], prec)
            assert is_approx(P[k] @ xs[k] - s[k], out["pi"][i, k], prec)
            assert is_approx(P[k], _unflat(out["P"][i, k], nx, nx), prec)
            assert is_approx(s[k], -out["p"][i, k], prec)
        for k in range(N):
            assert is_approx(us[k], out["u"][i, k], prec)
            assert is_approx(K[k], _unflat(out["K"][i, k], nu, nx), prec)
            assert is_approx(kf[k], out["k"][i, k], prec)


def _kkt_check(pkg, dims_d, arrays, out, i, tol):
    """Independent verification that (x,u,pi,lam,t) is a KKT point of QP i."""
    N, nx, nu, ng, nbx, nbu = (dims_d[k] for k in ("N", "nx", "nu", "ng", "nbx", "nbu"))
    ngN = dims_d["ngN"]
    A = _unflat(arrays["A"][i], nx, nx); Bm = _unflat(arrays["Bm"][i], nx, nu)
    Q = _unflat(arrays["Q"][i], nx, nx); S = _unflat(arrays["S"][i], nu, nx); R = _unflat(arrays["R"][i], nu, nu)
    x, u, pi, lam, t = out["x"][i], out["u"][i], out["pi"][i], out["lam"][i], out["t"][i]
    o = 0
    worst = 0.0
    for k in range(N + 1):
        nb = (nbu if k < N else 0) + (nbx if k > 0 else 0)
        ngk = ng if k < N else ngN
        nc = nb + ngk
        ll, lu = lam[o:o + nc], lam[o + nc:o + 2 * nc]
        tl, tu = t[o:o + nc], t[o + nc:o + 2 * nc]
        o += 2 * nc
        # constraint rows J [u;x]
        rows = []
        lo, up, ml, mu = [], [], [], []
        if k < N:
            for j in range(nbu):
                e = np.zeros(nu + nx); e[arrays["idxbu"][j]] = 1; rows.append(e)
                lo.append(arrays["lbu"][i, k, j]); up.append(arrays["ubu"][i, k, j]); ml.append(1); mu.append(1)
        if k > 0:
            for j in range(nbx):
                e = np.zeros(nu + nx); e[nu + arrays["idxbx"][j]] = 1; rows.append(e)
                lo.append(arrays["lbx"][i, k, j]); up.append(arrays["ubx"][i, k, j]); ml.append(1); mu.append(1)
        if k < N and ng:
            D = _unflat(arrays["D"][i, k], ng, nu)
            Cm = _unflat(arrays["C"][i, k], ng, nx) if k > 0 else np.zeros((ng, nx))
            for j in range(ng):
                rows.append(np.concatenate([D[j], Cm[j]]))
                lo.append(arrays["lg"][i, k, j]); up.append(arrays["ug"][i, k, j]); ml.append(1); mu.append(1)
        if k == N and ngN:
            Cm = _unflat(arrays["CN"][i], ngN, nx)
            for j in range(ngN):
                rows.append(np.concatenate([np.zeros(nu), Cm[j]]))
                lo.append(arrays["lgN"][i, j]); up.append(arrays["ugN"][i, j]); ml.append(1); mu.append(1)
        J = np.array(rows).reshape(nc, nu + nx)
        uk = u[k] if k < N else np.zeros(nu)
        z = np.concatenate([uk, x[k]])
        # stationarity wrt x_k (k>=1) and u_k
        if k >= 1:
            gx = Q[k] @ x[k] + arrays["q"][i, k] - pi[k]
            if k < N:
                gx += S[k].T @ u[k] + A[k].T @ pi[k + 1]
            gx += J[:, nu:].T @ (lu - ll)
            worst = max(worst, np.abs(gx).max())
        if k < N:
            gu = R[k] @ u[k] + S[k] @ x[k] + arrays["r"][i, k] + Bm[k].T @ pi[k + 1] + J[:, :nu].T @ (lu - ll)
            worst = max(worst, np.abs(gu).max())
            worst = max(worst, np.abs(A[k] @ x[k] + Bm[k] @ u[k] + arrays["b"][i, k] - x[k + 1]).max())
        if nc:
            v = J @ z
            # stage 0: the x-part of general rows is dropped by the embedding (C0 ignored)
            if k == 0:
                v = J[:, :nu] @ uk
            worst = max(worst, np.abs(v - np.array(lo) - tl).max(), np.abs(np.array(up) - v - tu).max())
            worst = max(worst, np.abs(ll * tl).max(), np.abs(lu * tu).max())
            assert (ll >= 0).all() and (lu >= 0).all() and (tl > 0).all() and (tu > 0).all()
    assert worst < tol, worst


@pytest.mark.parametrize("ric_alg", [0, 1])
def test_constrained_random(pkg, orc, ric_alg):
    """hpipm-cpp/test/ocp_qp_ipm_solver.cpp:112-168 (constrained): nx=5, nu=3, ng=2, box on u{0,1,2}
    and on two states: status Success and x[0] == x0; plus an independent KKT check."""
    B = 6
    dims_d, arrays = pkg.workload.random_qp(B, N=20, nx=5, nu=3, ng=2, nbx=2, nbu=3, seed=11, a_scale=0.4)
    # SPEED mode has no iterative refinement: stationarity stalls near 1e-9 on these badly scaled
    # problems, so use HPIPM's SPEED default res_g_max = 1e-6 for tol_stat
    args = orc.ipm_args(ric_alg=ric_alg, iter_max=40, tol_stat=1e-6)
    out = orc.qp_solve(_dims(pkg, dims_d), args, arrays, B)
    assert (out["status"] == 0).all(), out["status"]
    assert np.allclose(out["x"][:, 0], arrays["x0"])
    for i in range(B):
        _kkt_check(pkg, dims_d, arrays, out, i, 1e-6)


def test_ric_algs_agree(pkg, orc):
    B = 3
    dims_d, arrays = pkg.workload.random_qp(B, N=12, nx=6, nu=2, ng=3, nbx=1, nbu=2, seed=5, a_scale=0.4)
    o0 = orc.qp_solve(_dims(pkg, dims_d), orc.ipm_args(ric_alg=0, iter_max=40, tol_stat=1e-6), arrays, B)
    o1 = orc.qp_solve(_dims(pkg, dims_d), orc.ipm_args(ric_alg=1, iter_max=40, tol_stat=1e-6), arrays, B)
    assert (o0["iter"] == o1["iter"]).all()
    for k in ("x", "u", "pi", "lam"):
        assert np.allclose(o0[k], o1[k], rtol=1e-7, atol=1e-9)


def test_stat_table_layout(pkg, orc):
    """18-column statistics rows (ocp_qp_ipm_solver.cpp:381-403): row 0 = initial residuals, row i =
    iteration i; the last row's residuals equal the reported max residuals."""
    dims_d, arrays = pkg.workload.random_qp(1, N=8, nx=4, nu=2, nbu=2, seed=3)
    args = orc.ipm_args(iter_max=30)
    out = orc.qp_solve(_dims(pkg, dims_d), args, arrays, 1, stat_rows=32)
    it = int(out["iter"][0])
    assert it > 0
    st = out["stat"][0]
    assert np.allclose(st[it, 6:10], out["res_max"][0])
    assert (st[1:it + 1, 0] > 0).all() and (st[1:it + 1, 0] <= 1).all()  # alpha_aff
    assert (st[it + 1:] == 0).all()
    assert (np.diff(st[:it + 1, 5]) < 0).all()  # mu decreases monotonically here


def test_status_max_iter(pkg, orc):
    dims_d, arrays = pkg.workload.random_qp(1, N=8, nx=4, nu=2, nbu=2, seed=3)
    out = orc.qp_solve(_dims(pkg, dims_d), orc.ipm_args(iter_max=2), arrays, 1)
    assert out["status"][0] == 1 and out["iter"][0] == 2


def test_failed_pivot_zeroes_the_component(pkg, orc):
    """BLASFEO's potrf stores the inverse diagonal (0 for a non-positive pivot) and its trsv multiplies by it
    (blasfeo_common.h:71,76), so a failed pivot zeroes a component of the triangular solve instead of producing inf / nan.
    QP 1000454 of the all-stance N=50 workload hits such a pivot in its last iteration (mu ~ 2e-11): with a division by the
    zeroed pivot the oracle ended in NaN (status 3); with the inverse-diagonal semantics it converges in the 12 iterations
    the GPU takes (scripts/diag_qp.py 1000454 stance 50, profiles/r1_v14_parity_sweep_n50_stance.json)."""
    N = 50
    settings = dict(iter_max=50, alpha_min=1e-8, mu0=1e2, tol_stat=1e-6, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
                    reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
    w = pkg.workload.srbd_batch(1, N=N, contact_mode="stance", start=1000454)
    ref = orc.pipeline(orc.model_params(N), orc.ipm_args(**settings), N, pkg.capi.SRBD_HARD_INEQ, w["x"], w["u"],
                       w["xref"], w["x0"], w["contact"])
    assert ref["status"][0] == 0 and ref["iter"][0] == 12
    assert np.isfinite(ref["x"]).all() and np.isfinite(ref["u"]).all()
    assert ref["res_max"][0, 0] <= 1e-6 and (ref["res_max"][0, 1:] <= 1e-8).all()
