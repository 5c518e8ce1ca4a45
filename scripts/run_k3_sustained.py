"""K3 under SUSTAINED load (the board's 1000 W power cap bites after about a second): `reps` back-to-back solves of a
65536-QP batch, per-solve times from CUDA events; prints the mean of the second half and the SM clock seen by nvidia-smi."""
import sys, os, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import srbd_pkg
pkg = srbd_pkg.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
         reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
w = pkg.workload.srbd_batch(B, N=20, contact_mode="gait")
ctx = pkg.Context(B)
ctx.set_model(pkg.default_model_params(20)); ctx.set_ipm_args(pkg.default_ipm_args(**S))
ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
ctx.linearize(); ctx.assemble(1); ctx.qp_solve(); ctx.sync()
st = torch.cuda.ExternalStream(ctx.stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
with torch.cuda.stream(st):
    ev[0].record(st)
    for i in range(reps):
        ctx.qp_solve(); ev[i + 1].record(st)
time.sleep(0.05 * reps)
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
ctx.sync()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)])
half = ms[reps // 2:]
print("B=%d reps=%d  K3 first %.2f ms, second half mean %.2f ms -> %.0f solves/s sustained  [%s mid-run]  lib=%s" % (
    B, reps, ms[0], half.mean(), B / half.mean() * 1e3, clk, os.environ.get("SRBD_LIB", "in-tree")))
