"""BASELINE config 1 (the reference's controlLoop(): one problem, soft barrier, SQP + line search) on the GPU:
prints the SQP iteration count and the host->host time per NMPC solve; under `ncu --metrics gpu__time_duration.sum`
the launch list shows where the time goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg
pkg = srbd_pkg.load()
from srbd_nmpc_solver_b200.binding import make_dims
SOFT = pkg.capi.SRBD_BARRIER_SOFT
N = 20
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-4, tol_eq=1e-4, tol_ineq=1e-4, tol_comp=1e-4,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
w = pkg.workload.reference_nmpc_problem(N)
c = pkg.Context(1, make_dims(N=N))
c.set_model(pkg.default_model_params(N)); c.set_ipm_args(pkg.default_ipm_args(**S))


def nmpc():
    c.reset_sqp_state(); c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    for it in range(15):
        c.sqp_iterate(SOFT, do_line_search=True)
        if c.download_sqp_state()[1][0]:
            return it + 1
    return 15


reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nmpc()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); its = nmpc(); ts.append(time.perf_counter() - t0)
x, u = c.download_traj()
print("config 1: %d SQP iterations, %.3f ms per NMPC solve (median of %d), u0 = %s" % (
    its, 1e3 * float(np.median(ts)), reps, np.array2string(u[0, 0], precision=4)))
