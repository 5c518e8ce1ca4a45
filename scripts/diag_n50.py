import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg
pkg = srbd_pkg.load()
from oracle import oracle as orc
from srbd_nmpc_solver_b200.binding import make_dims
S = dict(iter_max=50, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
N, B = 50, 48
w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait", start=7)
for generic in ("0", "1"):
    os.environ["SRBD_K3_GENERIC"] = generic
    ctx = pkg.Context(B, make_dims(N=N)); ctx.set_model(pkg.default_model_params(N)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); ctx.sqp_iterate(1)
    st = ctx.download_stats(); lin, qp = ctx.download_linearization(), ctx.download_qp(); ctx.close()
    print("generic", generic, "status", st["status"].tolist()); print(" iter", st["iter"].tolist())
arrays = dict(A=lin["A"], Bm=lin["Bm"], b=lin["b"], Q=qp["Q"], S=qp["S"], R=qp["R"], q=qp["q"], r=qp["r"], D=qp["D"], lg=qp["lg"],
              ug=np.zeros_like(qp["lg"]), lg_mask=qp["lg_mask"], ug_mask=np.zeros_like(qp["lg"]), x0=w["x0"] - w["x"][:, 0])
ref = orc.qp_solve(make_dims(N=N), orc.ipm_args(**S), arrays, B, stat_rows=52, want=("x",))
print("oracle status", ref["status"].tolist()); print(" iter", ref["iter"].tolist())
bad = np.flatnonzero(ref["status"] != st["status"])
np.set_printoptions(linewidth=200, precision=3)
for i in bad[:2]:
    print("QP", i, "oracle stat tail:"); print(ref["stat"][i][max(0, ref["iter"][i]-6):ref["iter"][i]+1, :10])
    print(" gpu res_max", st["res_max"][i], "oracle res_max", ref["res_max"][i])
