"""Small cases through every kernel of the library (K1, K2 both instantiations, K3 variant + rescue stages + generic with
exports / refinement, K4, detection, MPC kernels, device SQP loop) for `compute-sanitizer --tool memcheck`."""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import srbd_pkg
pkg = srbd_pkg.load()
from srbd_nmpc_solver_b200.binding import make_dims
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8, reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
for B, N in ((13, 20), (3, 50), (7, 1), (160, 6)):   # 160 QPs > SMs: the throughput instantiation (compact BAbt streaming, lazy dense K1)
    w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait", spread=0.25)
    ctx = pkg.Context(B, make_dims(N=N)); ctx.set_model(pkg.default_model_params(N)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    ctx.sqp_iterate(1); st = ctx.download_stats(); print("variant", B, N, st["status"][:16], st["iter"][:16])
    lin, qp = ctx.download_linearization(), ctx.download_qp()     # lazy dense records
    ctx.sqp_iterate(1, do_line_search=True)
    it = ctx.sqp_solve(1, 3); print("sqp_solve", it)
    ctx.close()
# QP-level: detection + gated dispatch, generic kernel with exports + itref, MPC driver
dims_d, arrays, settings, A, Bm = pkg.workload.quadcopter_mpc()
ctx = pkg.Context(1, make_dims(**dims_d)); ctx.set_model(pkg.default_model_params(10)); ctx.set_ipm_args(pkg.default_ipm_args(**dict(settings, itref_corr_max=2)))
ctx.set_outputs(export_ric=True, export_stat=True)
ctx.qp_upload(arrays); ctx.qp_solve(); print("generic", ctx.download_stats()["iter"])
ctx.set_outputs(export_ric=False, export_stat=False)
xt, ut, it, st = ctx.mpc_run(A, Bm, np.zeros(12), np.zeros((1, 12)), 3); print("mpc", it.ravel())
ctx.close()
print("done")
