"""Tiny run of every kernel (K1, K2, K3 generic + SRBD variant, K4, pack) for compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg
pkg = srbd_pkg.load()
from srbd_nmpc_solver_b200.binding import make_dims
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
for N, B in ((20, 9), (3, 5)):
    w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait")
    for generic in ("0", "1"):
        os.environ["SRBD_K3_GENERIC"] = generic
        c = pkg.Context(B, make_dims(N=N)); c.set_model(pkg.default_model_params(N)); c.set_ipm_args(pkg.default_ipm_args(**S))
        c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
        c.sqp_iterate(1, do_line_search=True); c.sqp_iterate(0, do_line_search=True)
        st = c.download_stats(); c.close()
        print("srbd N", N, "generic", generic, st["iter"].tolist(), st["status"].tolist())
for shape, alg in ((dict(nx=5, nu=3, ng=2, nbx=2, nbu=3), 0), (dict(nx=6, nu=2, ng=3, nbx=1, nbu=2), 1), (dict(nx=12, nu=4, ng=0, nbx=3, nbu=4), 0)):
    d, a = pkg.workload.random_qp(3, N=6, seed=2, a_scale=0.4, **shape)
    c = pkg.Context(3, make_dims(**d)); c.set_ipm_args(pkg.default_ipm_args(**dict(S, ric_alg=alg, tol_stat=1e-6)))
    c.set_outputs(True, True); c.qp_upload(a); c.qp_solve(); st = c.download_stats(True); c.download_solution(); c.close()
    print("generic qp", shape, st["iter"].tolist(), st["status"].tolist())
print("done")
