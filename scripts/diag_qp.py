"""One QP of the seeded workload through every implementation: SRBD K3 variant, generic K3 kernel (with the per-iteration
statistics table) and the CPU oracle on the identical GPU-assembled data.  python scripts/diag_qp.py <qp index> [gait|stance]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg, numpy as np
pkg = srbd_pkg.load()
from oracle import oracle as orc
from srbd_nmpc_solver_b200.binding import make_dims
q = int(sys.argv[1]); contact = sys.argv[2] if len(sys.argv) > 2 else "gait"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 20
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8, reg_prim=1e-12,
         warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
if N > 20:
    S.update(iter_max=50, tol_stat=1e-6)
np.set_printoptions(linewidth=200, precision=3)
w = pkg.workload.srbd_batch(1, N=N, contact_mode=contact, start=q)
for generic in ("0", "1"):
    os.environ["SRBD_K3_GENERIC"] = generic
    ctx = pkg.Context(1, make_dims(N=N)); ctx.set_model(pkg.default_model_params(N)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
    if generic == "1":
        ctx.set_outputs(False, True)
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); ctx.linearize(); ctx.assemble(1); ctx.qp_solve(); ctx.sync()
    st = ctx.download_stats(generic == "1")
    print("K3 %s: iter %d status %d res_max %s" % ("generic" if generic == "1" else "SRBD variant", st["iter"][0], st["status"][0], st["res_max"][0]))
    if generic == "1":
        print(" it  a_aff    mu_aff   sigma    a_prim   a_dual   mu       res_stat res_eq   res_ineq res_comp")
        for i in range(min(int(st["iter"][0]) + 2, st["stat"].shape[1])):
            print("%3d " % i + " ".join("%8.2e" % v for v in st["stat"][0, i][:10]))
        lin, qp = ctx.download_linearization(), ctx.download_qp()
    ctx.close()
arrays = dict(A=lin["A"], Bm=lin["Bm"], b=lin["b"], Q=qp["Q"], S=qp["S"], R=qp["R"], q=qp["q"], r=qp["r"],
              D=qp["D"], lg=qp["lg"], ug=np.zeros_like(qp["lg"]), lg_mask=qp["lg_mask"],
              ug_mask=np.zeros_like(qp["lg"]), x0=w["x0"] - w["x"][:, 0])
ref = orc.qp_solve(make_dims(N=N), orc.ipm_args(**S), arrays, 1, stat_rows=52, want=("x", "u", "pi", "lam", "t"))
print("oracle (identical data): iter %d status %d res_max %s" % (ref["iter"][0], ref["status"][0], ref["res_max"][0]))
if "stat" in ref:
    for i in range(min(int(ref["iter"][0]) + 2, ref["stat"].shape[1])):
        print("%3d " % i + " ".join("%8.2e" % v for v in ref["stat"][0, i][:10]))
for k in ("x", "u", "pi", "lam", "t"):
    print("oracle", k, "has NaN:", bool(np.isnan(ref[k]).any()), " max |.| = %.3e" % np.nanmax(np.abs(ref[k])))
