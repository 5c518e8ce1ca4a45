"""One 65536-QP shard of the bench workload (config 3) by index: K3 time with CUDA events, status counts, the QPs the SRBD
variant alone does not converge on (SRBD_K3_NO_RESCUE=1).  Names the straggler rank of a multi-GPU run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srbd_pkg
pkg = srbd_pkg.load()
shard = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
w = pkg.workload.srbd_batch(B, N=20, contact_mode="gait", start=shard * B)
ctx = pkg.Context(B)
ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
ctx.linearize(); ctx.assemble(1); ctx.qp_solve(); ctx.sync()
st = torch.cuda.ExternalStream(ctx.stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
with torch.cuda.stream(st):
    ev[0].record(st)
    for i in range(4):
        ctx.qp_solve(); ev[i + 1].record(st)
ctx.sync()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
s = ctx.download_stats()
bad = np.flatnonzero(s["status"] != 0)
print("shard %d rescue=%s  K3 %.2f ms (min of 4)  status counts %s  iter max %d  not converged: %s" % (
    shard, os.environ.get("SRBD_K3_NO_RESCUE", "0") != "1", min(ms), np.bincount(s["status"], minlength=4).tolist(),
    s["iter"].max(), [(int(shard * B + i), int(s["status"][i]), int(s["iter"][i])) for i in bad[:8]]))
