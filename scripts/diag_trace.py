"""Per-iteration IPM statistics (the 18-column HPIPM stat table) of selected QPs through the generic K3 kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg, numpy as np
pkg = srbd_pkg.load()
qs = [int(a) for a in sys.argv[1:]]
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8, reg_prim=1e-12,
         warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
np.set_printoptions(linewidth=200, precision=3)
for q in qs:
    w = pkg.workload.srbd_batch(1, N=20, contact_mode="gait", start=q)
    ctx = pkg.Context(1); ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
    ctx.set_outputs(False, True)
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); ctx.linearize(); ctx.assemble(1); ctx.qp_solve(); ctx.sync()
    st = ctx.download_stats(True)
    it = int(st["iter"][0])
    print("QP", q, "iter", it, "status", int(st["status"][0]))
    print(" it  a_aff    mu_aff   sigma    a_prim   a_dual   mu       res_stat res_eq   res_ineq res_comp")
    for i in range(min(it + 2, st["stat"].shape[1])):
        r = st["stat"][0, i]
        print("%3d " % i + " ".join("%8.2e" % v for v in r[:10]))
    ctx.close()
