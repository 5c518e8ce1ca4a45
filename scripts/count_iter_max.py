"""How many QPs does the SRBD variant ALONE (SRBD_K3_NO_RESCUE=1) fail to converge on?  `shards` x 65536 QPs of the bench
workload (gait) or config 2 (stance).  Compares arithmetic variants of K3 (SRBD_LIB): every variant has its own knife-edge QPs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import srbd_pkg
pkg = srbd_pkg.load()
os.environ["SRBD_K3_NO_RESCUE"] = "1"
mode = sys.argv[1] if len(sys.argv) > 1 else "gait"
shards = int(sys.argv[2]) if len(sys.argv) > 2 else 8
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
B = 65536
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
ctx = pkg.Context(B)
ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
bad, hist, tot = [], np.zeros(64, dtype=np.int64), 0
for sh in range(first, first + shards):
    w = pkg.workload.srbd_batch(B, N=20, contact_mode=mode, start=sh * B)
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    ctx.sqp_iterate(1)
    s = ctx.download_stats()
    hist += np.bincount(s["iter"], minlength=64)[:64]; tot += B
    for i in np.flatnonzero(s["status"] != 0):
        bad.append((int(sh * B + i), int(s["status"][i]), int(s["iter"][i])))
print("%s %d QPs lib=%s: not converged %d %s  iterations>=17: %d  mean %.4f" % (
    mode, tot, os.environ.get("SRBD_LIB", "in-tree"), len(bad), bad[:12], int(hist[17:].sum()), float((hist * np.arange(64)).sum() / tot)))
