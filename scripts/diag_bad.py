"""Repeat a batch solve and list the QPs that did not converge (race / robustness diagnosis)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg, numpy as np
pkg = srbd_pkg.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8, reg_prim=1e-12,
         warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
START = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = pkg.workload.srbd_batch(B, N=20, contact_mode="gait", start=START)
ctx = pkg.Context(B); ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); ctx.linearize(); ctx.assemble(1)
ref = None
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    ctx.qp_solve(); ctx.sync()
    st = ctx.download_stats()
    bad = np.nonzero(st["status"] != 0)[0]
    it = st["iter"].copy()
    diff = 0 if ref is None else int((it != ref).sum())
    if ref is None: ref = it
    if diff: print("   differing QPs", np.nonzero(it != ref)[0][:10].tolist(), it[it != ref][:10].tolist(), ref[it != ref][:10].tolist())
    print("rep", rep, "bad", bad[:8].tolist(), "iters", st["iter"][bad[:8]].tolist(), "iter sum", int(it.sum()), "differs from rep0 in", diff, "QPs")
