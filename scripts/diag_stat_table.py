"""Statistics table of the tensor-core variant (exports on) and of the generic kernel against the oracle's, per column."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import srbd_pkg
pkg = srbd_pkg.load()
from srbd_nmpc_solver_b200.binding import make_dims
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle as orc
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
B, N = 96, 20
w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait", start=300)
def ctx_():
    c = pkg.Context(B, make_dims(N=N)); c.set_model(pkg.default_model_params(N)); c.set_ipm_args(pkg.default_ipm_args(**S)); return c
with ctx_() as ctx:
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); ctx.sqp_iterate(1)
    lin, qp = ctx.download_linearization(), ctx.download_qp()
arrays = dict(A=lin["A"], Bm=lin["Bm"], b=lin["b"], Q=qp["Q"], S=qp["S"], R=qp["R"], q=qp["q"], r=qp["r"],
              D=qp["D"], lg=qp["lg"], ug=np.zeros_like(qp["lg"]), lg_mask=qp["lg_mask"],
              ug_mask=np.zeros_like(qp["lg"]), x0=w["x0"] - w["x"][:, 0])
with ctx_() as ctx:
    ctx.qp_upload(arrays); ctx.set_outputs(export_ric=True, export_stat=True)
    ctx.qp_solve(); st_e = ctx.download_stats(with_table=True)
    os.environ["SRBD_K3_GENERIC"] = "1"; ctx.qp_solve(); st_g = ctx.download_stats(with_table=True)
ref = orc.qp_solve(make_dims(N=N), orc.ipm_args(**S), arrays, B, stat_rows=st_e["stat"].shape[1], want=("x",))
names = "alpha_aff mu_aff sigma alpha_prim alpha_dual mu res_stat res_eq res_ineq res_comp obj lq itp itc l0 l1 l2 l3".split()
for tag, st in (("variant", st_e), ("generic", st_g)):
    print(tag, "iters equal:", (st["iter"] == ref["iter"]).all())
    for c in range(18):
        a, b = st["stat"][:, :, c], ref["stat"][:, :, c]
        d = np.abs(a - b) / (np.abs(b) + 1e-12)
        i = np.unravel_index(np.argmax(d), d.shape)
        print("  %-10s max rel diff %.3e at qp %d row %d: gpu %.12e oracle %.12e (iter %d)" % (names[c], d.max(), i[0], i[1], a[i], b[i], ref["iter"][i[0]]))
