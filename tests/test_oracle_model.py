"""CPU tests (-m "not gpu") of the oracle's model part (dynamics/orientation_tool.h,
dynamics/SRBD_model.cpp, NMPC_solver.cpp:149-314).  The reference has no tests or fixtures for this
part ("parity unpinned"), so the oracle is checked by finite differences, by scipy's matrix
exponential, by an independently written numpy restatement, and by the SQP-level behaviour recorded
in SURVEY.md §4.4."""
import numpy as np
import pytest
from scipy.linalg import expm as sp_expm


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


RS = [np.array([0.3, -0.2, 0.1]), np.array([1e-3, 2e-3, -1e-3]), np.array([1.2, 0.4, -0.9]),
      np.array([0.0, 0.0, 0.7])]


@pytest.mark.parametrize("r", RS)
def test_so3_helpers(orc, r):
    R = orc.so3("expm", r)
    assert np.allclose(R, sp_expm(skew(r)), atol=1e-12)
    Jl, Jlt = orc.so3("jl", r), orc.so3("jlt", r)
    assert np.allclose(Jl @ Jlt, np.eye(3), atol=1e-9)
    # closed form of the left Jacobian: sum_k skew(r)^k/(k+1)!
    S, acc, term = skew(r), np.eye(3), np.eye(3)
    for k in range(1, 30):
        term = term @ S / (k + 1)
        acc = acc + term
    assert np.allclose(Jl, acc, atol=1e-9)
    dJ = orc.so3("djl", r)
    dJt = orc.so3("djlt", r)
    h = 1e-6
    for k in range(3):
        e = np.zeros(3); e[k] = h
        fd = (orc.so3("jl", r + e) - orc.so3("jl", r - e)) / (2 * h)
        assert np.allclose(dJ[k], fd, atol=2e-8)
        fdt = (orc.so3("jlt", r + e) - orc.so3("jlt", r - e)) / (2 * h)
        assert np.allclose(dJt[k], fdt, atol=2e-8)
    assert np.allclose(orc.so3("skew", r), skew(r))


def test_so3_small_angle_clamp(orc):
    """theta is clamped at 1e-10 (orientation_tool.h:78-83): r = 0 must give finite identity-like results."""
    z = np.zeros(3)
    assert np.allclose(orc.so3("expm", z), np.eye(3))
    assert np.allclose(orc.so3("jl", z), np.eye(3), atol=1e-9)
    assert np.allclose(orc.so3("jlt", z), np.eye(3), atol=1e-9)
    assert np.isfinite(orc.so3("djlt", z)).all()


def np_continuous(m, x, u):
    """Independent numpy restatement of SRBD_model.cpp:75-98 (values only)."""
    r, l, p, v = x[0:3], x[3:6], x[6:9], x[9:12]
    th = max(np.linalg.norm(r), 1e-10)
    V = skew(r) / th
    R = sp_expm(skew(r))
    Iinv = np.array(m.inertia_inv).reshape(3, 3).T
    w = R @ Iinv @ R.T @ l
    cot = 1.0 / np.tan(0.5 * th)
    Jlt = 0.5 * cot * th * np.eye(3) + (1 - 0.5 * cot * th) * (V @ V + np.eye(3)) - 0.5 * th * V
    pf = np.array(m.foot_pos)
    FR, TR, FL, TL = u[0:3], u[3:6], u[6:9], u[9:12]
    dx = np.zeros(12)
    dx[0:3] = Jlt @ w
    dx[3:6] = TR + TL + np.cross(pf[0:3] - p, FR) + np.cross(pf[3:6] - p, FL)
    dx[6:9] = v
    dx[9:12] = (FR + FL) / m.mass + np.array(m.gravity)
    return dx


def test_continuous_dynamics_and_jacobians(orc):
    m = orc.model_params(20)
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = rng.uniform(-0.5, 0.5, 12); x[8] += 1.0
        u = rng.uniform(-50, 50, 12); u[2] += 80; u[8] += 80
        dx, jfx, jfu = orc.continuous(m, x, u)
        assert np.allclose(dx, np_continuous(m, x, u), rtol=1e-11, atol=1e-11)
        h = 1e-6
        for j in range(12):
            e = np.zeros(12); e[j] = h
            fdx = (orc.continuous(m, x + e, u, jac=False) - orc.continuous(m, x - e, u, jac=False)) / (2 * h)
            assert np.allclose(jfx[:, j], fdx, atol=5e-7 * max(1.0, np.abs(fdx).max())), j
            fdu = (orc.continuous(m, x, u + e, jac=False) - orc.continuous(m, x, u - e, jac=False)) / (2 * h)
            assert np.allclose(jfu[:, j], fdu, atol=1e-7), j


def test_reference_initial_guess_values(orc):
    """at x=0 (p_z=1), u=100: xdot = (0,0,0, 400,0,200, 0,0,0, 13.33,13.33,3.533),
    nnz(df/dx) = 12, nnz(df/du) = 20."""
    m = orc.model_params(20)
    x = np.zeros(12); x[8] = 1.0
    u = np.full(12, 100.0)
    dx, jfx, jfu = orc.continuous(m, x, u)
    assert np.allclose(dx, [0, 0, 0, 400, 0, 200, 0, 0, 0, 200 / 15, 200 / 15, 200 / 15 - 9.8], atol=1e-9)
    assert np.count_nonzero(np.abs(jfx) > 1e-14) == 12
    assert np.count_nonzero(jfu) == 20


def test_shooting(orc):
    """SRBD_model.cpp:143-235: f = x_next - RK4(x,u); A = I + dt*jfx(x,u); B = dt*jfu(x); b = -f."""
    m = orc.model_params(20)
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.3, 0.3, 12); x[8] += 1.0
    xn = x + rng.uniform(-0.05, 0.05, 12)
    u = rng.uniform(-20, 20, 12); u[2] += 70; u[8] += 70
    A, B, b, f = orc.shooting(m, x, xn, u)
    dt = m.dt
    k1 = np_continuous(m, x, u)
    k2 = np_continuous(m, x + 0.5 * dt * k1, u)
    k3 = np_continuous(m, x + 0.5 * dt * k2, u)
    k4 = np_continuous(m, x + dt * k3, u)
    xg = x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    assert np.allclose(f, xn - xg, atol=1e-12)
    assert np.allclose(b, -f)
    _, jfx, jfu = orc.continuous(m, x, u)
    assert np.allclose(A, np.eye(12) + dt * jfx, atol=1e-15)
    assert np.allclose(B, dt * jfu, atol=1e-15)


def test_constraint_rows_and_barrier(orc):
    """SRBD_model.cpp:237-295."""
    m = orc.model_params(20)
    u = np.array([3., -2., 50., 0.4, -0.3, 0.2, -1., 4., 60., 0.1, 0.2, -0.5])
    Ac, f = orc.constraint(m, u)
    mu, L = 0.5, 0.05
    for leg in range(2):
        F, T = u[6 * leg:6 * leg + 3], u[6 * leg + 3:6 * leg + 6]
        exp = [-F[0] + mu * F[2], -F[1] + mu * F[2], F[0] + mu * F[2], F[1] + mu * F[2], -F[2] + 1000.0, F[2] - 0.0,
               L * F[2] - T[1], L * F[2] + T[1], L * F[2] - T[2], L * F[2] + T[2], -T[0], T[0]]
        assert np.allclose(f[12 * leg:12 * leg + 12], exp)
    assert np.allclose(Ac @ u + np.where(np.arange(24) % 12 == 4, 1000.0, 0.0), f)
    # swing contact (extension): fmax := swing_fmax = 1
    _, fs = orc.constraint(m, u, stance=[0, 1])
    assert np.isclose(fs[4], -u[2] + 1.0) and np.isclose(fs[16], -u[8] + 1000.0)
    # barrier: C1-continuity at theta and derivative checks
    for v in (0.3, 4.999, 5.001, 40.0, -3.0):
        b, db, ddb = orc.barrier(v, 0.1, 5.0)
        h = 1e-6
        bp, bm = orc.barrier(v + h, 0.1, 5.0)[0], orc.barrier(v - h, 0.1, 5.0)[0]
        assert np.isclose(db, (bp - bm) / (2 * h), atol=1e-7)
        assert ddb > 0
    lo, hi = orc.barrier(5.0, 0.1, 5.0), orc.barrier(5.0 + 1e-12, 0.1, 5.0)
    assert np.allclose(lo, hi, atol=1e-10)


def test_assemble_soft_matches_formulae(pkg, orc):
    """NMPC_solver.cpp:297-313."""
    m = orc.model_params(20)
    N = 6
    w = pkg.workload.srbd_batch(3, N=N, contact_mode="gait")
    o = orc.assemble(m, N, 0, w["x"], w["u"], w["xref"], w["contact"])
    Qd, Qf = np.array(m.Q), np.array(m.Qf)
    for i in range(3):
        for k in range(N):
            Ac, f = orc.constraint(m, w["u"][i, k], stance=w["contact"][i, k])
            bar = np.array([orc.barrier(v, m.mu_b, m.theta_b) for v in f])
            R = m.R * np.eye(12) + Ac.T @ np.diag(bar[:, 2]) @ Ac
            r = m.R * w["u"][i, k] + Ac.T @ bar[:, 1]
            assert np.allclose(o["R"][i, k].reshape(12, 12).T, R, rtol=1e-13, atol=1e-15)
            assert np.allclose(o["r"][i, k], r, rtol=1e-13, atol=1e-15)
            assert np.allclose(o["q"][i, k], Qd * (w["x"][i, k] - w["xref"][i, k]))
            assert np.allclose(o["Q"][i, k].reshape(12, 12), np.diag(Qd))
            assert (o["S"][i, k] == 0).all()
            A, B, b, fd = orc.shooting(m, w["x"][i, k], w["x"][i, k + 1], w["u"][i, k])
            assert np.array_equal(o["A"][i, k].reshape(12, 12).T, A)
            assert np.array_equal(o["b"][i, k], b) and np.array_equal(o["defect"][i, k], fd)
        assert np.allclose(o["Q"][i, N].reshape(12, 12), np.diag(Qf))
        assert np.allclose(o["q"][i, N], Qf * (w["x"][i, N] - w["xref"][i, N]))


def test_assemble_hard(pkg, orc):
    """HARD_INEQ (the variant commented out at NMPC_solver.cpp:300-304): D = Ac, lg = -(Ac u + bc), the
    +-x^T tau rows stay a relaxed barrier (20 = 24-4 hard rows, cf. the 20-row C at :301)."""
    m = orc.model_params(20)
    N = 4
    w = pkg.workload.srbd_batch(2, N=N, contact_mode="gait")
    o = orc.assemble(m, N, 1, w["x"], w["u"], w["xref"], w["contact"])
    soft = np.isin(np.arange(24) % 12, (10, 11))
    for i in range(2):
        for k in range(N):
            Ac, f = orc.constraint(m, w["u"][i, k], stance=w["contact"][i, k])
            assert np.allclose(o["D"][i, k].reshape(12, 24).T, Ac)
            assert np.allclose(o["lg"][i, k], -f)
            assert np.array_equal(o["lg_mask"][i, k], (~soft).astype(float))
            bar = np.array([orc.barrier(v, m.mu_b, m.theta_b) for v in f])
            bar[~soft] = 0
            assert np.allclose(o["R"][i, k].reshape(12, 12).T, m.R * np.eye(12) + Ac.T @ np.diag(bar[:, 2]) @ Ac)
            assert np.allclose(o["r"][i, k], m.R * w["u"][i, k] + Ac.T @ bar[:, 1])


def test_reference_sqp_loop(pkg, orc):
    """Config 1 = the reference's controlLoop() (NMPC_solver.cpp:353-380) restated with the oracle:
    SURVEY.md §4.4 recorded (from an independent numpy restatement) convergence after 11 SQP
    iterations, alpha = 1 for iterations 0-5 then 0.5, and the converged u0."""
    N = 20
    m = orc.model_params(N)
    w = pkg.workload.reference_nmpc_problem(N)
    x, u, xref, x0 = w["x"][0].copy(), w["u"][0].copy(), w["xref"][0], w["x0"][0]
    args = orc.ipm_args(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-4, tol_eq=1e-4, tol_ineq=1e-4,
                        tol_comp=1e-4, reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
    alpha, alphas, conv_it = 1.0, [], None
    for it in range(15):
        out = orc.pipeline(m, args, N, 0, x[None], u[None], xref[None], x0[None])
        assert out["iter"][0] == 0 and out["status"][0] == 0  # soft barrier => unconstrained QP, iter == 0
        x, u, alpha, conv, merit = orc.line_search(m, N, x, u, xref, out["x"][0], out["u"][0], alpha)
        alphas.append(alpha)
        if conv:
            conv_it = it + 1
            break
    assert conv_it == 11
    assert alphas[:6] == [1.0] * 6 and alphas[6:] == [0.5] * 5
    u0_expected = [54.37, 48.28, 100.32, 4.46, 24.94, 5.55, 63.53, 59.05, 122.50, 4.46, 25.97, 6.28]
    assert np.allclose(u[0], u0_expected, atol=0.02)


def test_line_search_merit_by_assemble_mode(pkg, orc):
    """orc_line_search_mode: in SRBD_HARD_INEQ only the rows the assembly keeps as a relaxed barrier (+-x^T tau of each
    contact) enter phi / dphi -- the hard rows are constraints of the QP, not cost terms; mode 0 is the reference
    (NMPC_solver.cpp:166-187: all 24 rows).  The constraint violation theta does not depend on the mode."""
    N = 6
    w = pkg.workload.srbd_batch(1, N=N, contact_mode="gait", spread=0.25)
    m = orc.model_params(N)
    rng = np.random.default_rng(2)
    dx, du = 1e-3 * rng.standard_normal((N + 1, 12)), 1e-2 * rng.standard_normal((N, 12))
    outs = [orc.line_search(m, N, w["x"][0], w["u"][0], w["xref"][0], dx, du, 1.0, contact=w["contact"][0], mode=md)
            for md in (0, 1)]
    (_, _, _, _, m0), (_, _, _, _, m1) = outs
    assert m0[2] == m1[2] and m0[2] > 0.0                   # theta
    assert m0[0] != m1[0] and m0[1] != m1[1]                # phi, dphi
    # the difference of phi is the barrier of the 20 hard rows: -mu_b log(v) for v > theta_b etc.: recompute it
    diff = 0.0
    for k in range(N):
        Ac, f = orc.constraint(m, w["u"][0, k], stance=w["contact"][0, k])
        for g in range(24):
            if g % 12 in (10, 11):
                continue
            diff += orc.barrier(float(f[g]), m.mu_b, m.theta_b)[0]
    assert abs((m0[0] - m1[0]) - diff) <= 1e-9 * max(1.0, abs(diff))
