"""Dumps the K3 outputs (x, u, pi, lam, t, iter) of a seeded batch to an .npz (SRBD_LIB selects the build);
`python scripts/ab_k3_bits.py cmp a.npz b.npz` compares two dumps bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if sys.argv[1] == "cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = [k for k in a.files if not np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8))]
    print("bitwise identical (%d arrays)" % len(a.files) if not bad else "DIFFERENT: %s" % bad)
    for k in bad:
        d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
        print("  ", k, "max abs diff", d.max(), "entries", int((d > 0).sum()))
    sys.exit(1 if bad else 0)
import srbd_pkg
pkg = srbd_pkg.load()
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
w = pkg.workload.srbd_batch(B, N=20, contact_mode="gait", start=500000)
ctx = pkg.Context(B)
ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
ctx.sqp_iterate(1)
sol = ctx.download_solution(want=("x", "u", "pi", "lam", "t"))
st = ctx.download_stats()
np.savez(sys.argv[1], iter=st["iter"], status=st["status"], res_max=st["res_max"], **sol)
print("dumped", sys.argv[1], "iter mean %.4f" % st["iter"].mean())
