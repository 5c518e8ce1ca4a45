"""Diagnostic: where do GPU-vs-oracle differences in lam come from?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg
pkg = srbd_pkg.load()
from oracle import oracle as orc
from srbd_nmpc_solver_b200.binding import make_dims
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
B, N = 512, 20
def rel(a, b):
    a, b = a.reshape(B, -1), b.reshape(B, -1)
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)
for contact in ("stance", "gait"):
    w = pkg.workload.srbd_batch(B, N=N, contact_mode=contact)
    ctx = pkg.Context(B)
    ctx.set_model(pkg.default_model_params(N)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    ctx.sqp_iterate(1)
    sol = ctx.download_solution(want=("x", "u", "pi", "lam", "t")); st = ctx.download_stats()
    lin = ctx.download_linearization(); qp = ctx.download_qp()
    ctx.close()
    ref = orc.pipeline(orc.model_params(N), orc.ipm_args(**S), N, 1, w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    # oracle IPM on the GPU-assembled QP data (identical inputs)
    arrays = dict(A=lin["A"], Bm=lin["Bm"], b=lin["b"], Q=qp["Q"], S=qp["S"], R=qp["R"], q=qp["q"], r=qp["r"],
                  D=qp["D"], lg=qp["lg"], ug=np.zeros_like(qp["lg"]), lg_mask=qp["lg_mask"], ug_mask=np.zeros_like(qp["lg"]),
                  x0=w["x0"] - w["x"][:, 0])
    ref2 = orc.qp_solve(make_dims(N=N), orc.ipm_args(**S), arrays, B, want=("x", "u", "pi", "lam", "t"))
    print("==", contact, "iters equal:", (st["iter"] == ref["iter"]).all(), (st["iter"] == ref2["iter"]).all())
    for name, r in (("oracle pipeline (own K1/K2)", ref), ("oracle IPM on GPU QP data", ref2)):
        print(" ", name)
        for k in ("x", "u", "lam", "t"):
            e = rel(sol[k], r[k])
            print("    %-4s max %.2e  p99 %.2e  median %.2e  frac<=1e-9 %.4f" % (k, e.max(), np.quantile(e, 0.99), np.median(e), (e <= 1e-9).mean()))
        e = rel(sol["pi"][:, 1:], r["pi"][:, 1:]); print("    pi   max %.2e median %.2e" % (e.max(), np.median(e)))
    # worst QP: which entries
    e = rel(sol["lam"], ref2["lam"]); i = int(np.argmax(e))
    d = np.abs(sol["lam"][i] - ref2["lam"][i]); j = np.argsort(-d)[:5]
    print("  worst QP", i, "iter", st["iter"][i], "entries", j, "lam", ref2["lam"][i][j], "dlam", d[j], "t", ref2["t"][i][j])
    # oracle-vs-oracle sensitivity: perturb the QP data by 1 ulp and re-solve
    arr3 = dict(arrays); arr3["r"] = np.nextafter(arrays["r"], np.inf)
    ref3 = orc.qp_solve(make_dims(N=N), orc.ipm_args(**S), arr3, B, want=("x", "u", "pi", "lam", "t"))
    for k in ("x", "u", "lam", "t"):
        e = rel(ref3[k], ref2[k]); print("    oracle self-sensitivity (r + 1ulp) %-4s max %.2e median %.2e" % (k, e.max(), np.median(e)))
    print("    iters equal under 1ulp perturbation:", (ref3["iter"] == ref2["iter"]).all())
