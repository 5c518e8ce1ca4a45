"""The whole throughput path WITH the rescue pass at scale: `shards` x 65536 QPs of the bench workload (gait) and of config 2
(stance), default settings of the library (compact BAbt streaming, lazy dense records: the rescue list gets its dense
records from K1 / K2 runs over the list).  Reports the QPs that did not converge (expected: none) and, for every shard on
which the variant alone would have left a QP at iter_max, that the SRBD_K1_DENSE=1 / SRBD_K2_DENSE=1 run gives the same
statuses and iteration counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import srbd_pkg
pkg = srbd_pkg.load()
shards = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = 65536
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)


def solve(w, env):
    for k in ("SRBD_K1_DENSE", "SRBD_K2_DENSE", "SRBD_K3_NO_RESCUE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = pkg.Context(B)
    ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    ctx.sqp_iterate(1)
    s = ctx.download_stats()
    ctx.close()
    return s


for mode in ("gait", "stance"):
    tot = bad = rescued = 0
    for sh in range(shards):
        w = pkg.workload.srbd_batch(B, N=20, contact_mode=mode, start=sh * B)
        s = solve(w, {})
        raw = solve(w, {"SRBD_K3_NO_RESCUE": "1"})
        tot += B; bad += int((s["status"] != 0).sum())
        hit = np.flatnonzero(raw["status"] != 0)
        if len(hit):
            d = solve(w, {"SRBD_K1_DENSE": "1", "SRBD_K2_DENSE": "1"})
            same = np.array_equal(d["status"], s["status"]) and np.array_equal(d["iter"], s["iter"])
            rescued += len(hit)
            print("  %s shard %d: variant alone leaves %s at status %s after %s iterations; with the rescue: status %s, %s iterations; "
                  "lazy dense records == dense records: %s" % (mode, sh, (sh * B + hit).tolist(), raw["status"][hit].tolist(),
                                                              raw["iter"][hit].tolist(), s["status"][hit].tolist(), s["iter"][hit].tolist(), same))
    print("%s: %d QPs, not converged with the rescue: %d, QPs the rescue took over: %d" % (mode, tot, bad, rescued))
