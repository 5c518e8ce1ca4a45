"""Single-instance timing of K3 (BASELINE config 5 shape: one QP, N = 50, all stance): median of `reps` solves of ONE QP,
device-resident (K3 only) -- isolates the kernel from the host path.  SRBD_LIB selects the build."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import srbd_pkg
pkg = srbd_pkg.load()
from srbd_nmpc_solver_b200.binding import make_dims
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
S = dict(iter_max=50, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
w = pkg.workload.srbd_batch(B, N=N, contact_mode="stance")
ctx = pkg.Context(B, make_dims(N=N))
ctx.set_model(pkg.default_model_params(N)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
ctx.linearize(); ctx.assemble(1); ctx.qp_solve(); ctx.sync()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); ctx.qp_solve(); ctx.sync(); ts.append(time.perf_counter() - t0)
st = ctx.download_stats()
if os.environ.get("SRBD_PROF"):   # a -DSRBD_K3_PROFILE=1 build: per-sweep clock64 sums of the last solve (leader warp)
    h = ctx.batch_stats()["iter_hist"][48:53]
    it = float(st["iter"].sum()) * N
    print("cycles per stage and iteration: residual %.0f  factor %.0f  backvec %.0f  forward(pred) %.0f  forward(fin) %.0f  | total %.0f" % (
        *[v / it for v in h], sum(h) / it))
print("B=%d N=%d iters=%s status=%s  K3 p50 %.1f us  p99 %.1f us  lib=%s" % (
    B, N, st["iter"][:4], st["status"][:4], 1e6 * np.percentile(ts, 50), 1e6 * np.percentile(ts, 99), os.environ.get("SRBD_LIB", "in-tree")))
