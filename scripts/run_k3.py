"""Small driver for profiling / timing: passes of the hot path on B QPs (default one resident wave)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg
pkg = srbd_pkg.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
w = pkg.workload.srbd_batch(B, N=20, contact_mode="gait")
ctx = pkg.Context(B)
ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
ctx.linearize(); ctx.assemble(1); ctx.qp_solve(); ctx.sync()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); ctx.qp_solve(); ctx.sync(); ts.append(time.perf_counter() - t0)
st = ctx.download_stats()
if os.environ.get("SRBD_PROF"):   # a -DSRBD_K3_PROFILE=1 build: per-sweep clock64 sums of the last solve
    h = ctx.batch_stats()["iter_hist"][48:53]
    tot = float(sum(h))
    print("per-sweep cycles per QP: residual %.0f  factor %.0f  backvec %.0f  forward(pred) %.0f  forward(fin) %.0f  | shares %s" % (
        *[v / B for v in h], " ".join("%.1f%%" % (100 * v / tot) for v in h)))
t = min(ts)
print("ok B=%d iters=%.3f allconv=%s  K3 %.2f ms  -> %.0f solves/s (generic=%s)" % (
    B, st["iter"].mean(), (st["status"] == 0).all(), 1e3 * t, B / t, os.environ.get("SRBD_K3_GENERIC", "0")))
