import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import srbd_pkg
pkg = srbd_pkg.load()
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
for mode, qp in (("gait", 318147), ("stance", 1007927), ("gait", 131072+18573)):
    w = pkg.workload.srbd_batch(8, N=20, contact_mode=mode, start=qp - 3)
    ctx = pkg.Context(8)
    ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    ctx.sqp_iterate(1); s = ctx.download_stats()
    print(mode, qp, "status", s["status"].tolist(), "iter", s["iter"].tolist(), "res", s["res_max"][3].tolist(), "lib", os.environ.get("SRBD_LIB","in-tree"), "norescue", os.environ.get("SRBD_K3_NO_RESCUE"))
    ctx.close()
