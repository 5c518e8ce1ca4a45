"""CPU tests (-m "not gpu") of the higher-precision ARBITER: oracle/ocp_qp_ipm.c compiled with __float128 arithmetic
(oracle/Makefile, -DORC_QUAD; entry points orc_qp_solve_*_q).  Same algorithm, constants, operation order and status
logic as the double oracle; where the GPU and the double oracle differ it says which side is closer to the
exact-arithmetic iterates (tests/test_gpu_parity.py::check_iterates, scripts/parity_sweep.py)."""
import numpy as np


def relerr(a, b):
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)


def test_arbiter_reproduces_the_golden_vectors(pkg, orc, golden_quadcopter):
    """hpipm-cpp/test/ocp_qp_ipm_solver.cpp:170-315 in __float128: the reference's sol0..14.txt at isApprox(1e-9) and
    the iteration counts of the double oracle -- the arbiter is pinned by the same vectors."""
    from srbd_nmpc_solver_b200.binding import make_dims
    dims_d, arrays, settings, A, Bm = pkg.workload.quadcopter_mpc()
    dims = make_dims(**dims_d)
    x, iters = np.zeros(12), []
    for t in range(15):
        arrays["x0"] = x[None, :].copy()
        out = orc.qp_solve(dims, orc.ipm_args(**settings), arrays, 1, quad=True)
        assert out["status"][0] == 0
        cat = np.concatenate([out["x"][0].reshape(-1), out["u"][0].reshape(-1)])
        g = golden_quadcopter[t]
        assert np.linalg.norm(cat - g) <= 1e-9 * min(np.linalg.norm(cat), np.linalg.norm(g)), t
        iters.append(int(out["iter"][0]))
        arrays["x_init"], arrays["u_init"] = out["x"].copy(), out["u"].copy()
        x = A @ x + Bm @ out["u"][0, 0]
    assert iters == [17, 10, 8, 7, 6, 7, 5, 5, 4, 3, 4, 4, 4, 3, 3]


def test_double_oracle_against_the_arbiter_on_srbd_qps(pkg, orc):
    """Config 3 QPs (N = 20, gait, HARD_INEQ, tol 1e-8): the double oracle takes the arbiter's iteration count on every
    QP and its x, t, pi are within 1e-9 of the exact-arithmetic iterates; u and above all lam are NOT on every QP --
    the rounding floor of the algorithm in double precision, which no double implementation,
    HPIPM included, can be expected to beat."""
    from srbd_nmpc_solver_b200.binding import make_dims
    B, N = 48, 20
    settings = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
                    reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
    w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait")
    o = orc.assemble(orc.model_params(N), N, 1, w["x"], w["u"], w["xref"], w["contact"])
    arrays = dict(A=o["A"], Bm=o["Bm"], b=o["b"], Q=o["Q"], S=o["S"], R=o["R"], q=o["q"], r=o["r"], D=o["D"],
                  lg=o["lg"], ug=np.zeros_like(o["lg"]), lg_mask=o["lg_mask"], ug_mask=np.zeros_like(o["lg"]),
                  x0=w["x0"] - w["x"][:, 0])
    want = ("x", "u", "pi", "lam", "t")
    d = orc.qp_solve(make_dims(N=N), orc.ipm_args(**settings), arrays, B, want=want)
    q = orc.qp_solve(make_dims(N=N), orc.ipm_args(**settings), arrays, B, want=want, quad=True)
    assert (d["status"] == 0).all() and (q["status"] == 0).all()
    assert (d["iter"] == q["iter"]).all()
    for k in ("x", "t"):
        assert relerr(d[k], q[k]).max() <= 1e-9, k
    assert relerr(d["pi"][:, 1:], q["pi"][:, 1:]).max() <= 1e-9
    assert relerr(d["u"], q["u"]).max() <= 1e-8
    e = relerr(d["lam"], q["lam"])
    assert (e <= 1e-9).mean() >= 0.5 and e.max() <= 1e-4   # degenerate active sets: multipliers are not unique
