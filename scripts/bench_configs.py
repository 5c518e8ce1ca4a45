"""Measures BASELINE.json configs 1, 2, 4, 5 (config 3 is bench.py's headline line).  Writes one JSON object."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srbd_pkg
pkg = srbd_pkg.load()
from oracle import oracle as orc
from srbd_nmpc_solver_b200.binding import make_dims
HARD, SOFT = pkg.capi.SRBD_HARD_INEQ, pkg.capi.SRBD_BARRIER_SOFT
S8 = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
          reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
out = {}


def ctx_for(B, N, settings):
    c = pkg.Context(B, make_dims(N=N))
    c.set_model(pkg.default_model_params(N)); c.set_ipm_args(pkg.default_ipm_args(**settings))
    return c


# ---- config 1: the reference's controlLoop() (one problem, soft barrier, SQP + line search) ----------------
N = 20
ref_set = dict(S8, tol_stat=1e-4, tol_eq=1e-4, tol_ineq=1e-4, tol_comp=1e-4)
w = pkg.workload.reference_nmpc_problem(N)
c = ctx_for(1, N, ref_set)
def nmpc_gpu():
    c.reset_sqp_state(); c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    for it in range(15):
        c.sqp_iterate(SOFT, do_line_search=True)
        if c.download_sqp_state()[1][0]:
            return it + 1
    return 15
nmpc_gpu()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); its = nmpc_gpu(); ts.append(time.perf_counter() - t0)
m = orc.model_params(N)
def nmpc_cpu():
    x, u, al = w["x"][0].copy(), w["u"][0].copy(), 1.0
    for it in range(15):
        o = orc.pipeline(m, orc.ipm_args(**ref_set), N, SOFT, x[None], u[None], w["xref"], w["x0"], threads=1)
        x, u, al, conv, _ = orc.line_search(m, N, x, u, w["xref"][0], o["x"][0], o["u"][0], al)
        if conv:
            return it + 1
    return 15
tc = []
for _ in range(5):
    t0 = time.perf_counter(); itc = nmpc_cpu(); tc.append(time.perf_counter() - t0)
out["config1_reference_workload"] = dict(sqp_iterations_gpu=its, sqp_iterations_cpu_oracle=itc,
    gpu_ms_per_nmpc_solve_cold=1e3 * float(np.median(ts)), cpu_oracle_ms_per_nmpc_solve_1core=1e3 * float(np.median(tc)),
    note="first (cold) repetition of controlLoop(): 11 SQP iterations; host->host incl. per-iteration convergence readback")
c.close()

# ---- config 2: B=1024, N=20, all contacts in stance ------------------------------------------------------------
B = 1024
w = pkg.workload.srbd_batch(B, N=20, contact_mode="stance")
c = ctx_for(B, 20, S8)
c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
for _ in range(3): c.sqp_iterate(HARD)
c.sync(); ts = []
for _ in range(10):
    t0 = time.perf_counter(); c.sqp_iterate(HARD); c.sync(); ts.append(time.perf_counter() - t0)
st = c.download_stats()
t0 = time.perf_counter(); o = orc.pipeline(orc.model_params(20), orc.ipm_args(**S8), 20, HARD, w["x"], w["u"], w["xref"], w["x0"], w["contact"], duals=False); tcpu = time.perf_counter() - t0
out["config2_b1024_stance"] = dict(gpu_solves_per_s=B / float(np.median(ts)), gpu_ms=1e3 * float(np.median(ts)), all_converged=bool((st["status"] == 0).all()),
    iter_mean=float(st["iter"].mean()), iteration_counts_equal_oracle=bool((st["iter"] == o["iter"]).all()),
    cpu_oracle_solves_per_s=B / tcpu, cpu_threads=orc.num_threads(),
    note="1024 QPs fill only 58% of one resident wave (1776 warps): latency-, not throughput-limited")
c.close()

# ---- config 4: B=4096, N=100, full SQP (K1+K2+K3+K4, <= 15 iterations) ---------------------------------------------
# Two start distributions: spread = 1.0 is the config-3 generator taken literally (CoM up to 0.28 m beside the feet: over a
# 1.5 s horizon no admissible wrench holds the body, the SQP linearisation is meaningless after a few stages) and
# spread = 0.25 a physically sensible one (workload.srbd_batch).  Two step-length policies: "carried" = the reference
# (alpha_ is a member that is never reset, NMPC_solver.h:104: once halved it stays halved) and "reset" = alpha := 1 at
# every SQP iteration (srbd_reset_sqp_state).
B, N = 4096, 100
S4 = dict(S8, iter_max=50, tol_stat=1e-6)
out["config4_b4096_n100_full_sqp"] = {}
for spread in (1.0, 0.25):
    for policy in ("carried", "reset"):
        w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait", spread=spread)
        c = ctx_for(B, N, S4)
        c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); c.sqp_iterate(HARD); c.sync()
        c.reset_sqp_state(); c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
        t0 = time.perf_counter(); hist = np.zeros(64, dtype=np.int64); nconv = []; qp_solves = 0; statuses = np.zeros(5, dtype=np.int64)
        ever = np.zeros(B, dtype=bool)
        for it in range(15):
            if policy == "reset":
                c.reset_sqp_state()
            c.sqp_iterate(HARD, do_line_search=True)
            bs = c.batch_stats(); hist += np.array(bs["iter_hist"]); statuses += np.array(bs["status_count"]); qp_solves += B
            conv = c.download_sqp_state()[1]; ever |= conv.astype(bool); nconv.append(int(conv.sum()))
        c.sync(); tt = time.perf_counter() - t0
        out["config4_b4096_n100_full_sqp"]["spread_%g_alpha_%s" % (spread, policy)] = dict(
            sqp_iterations=15, seconds=tt, qp_solves_per_s=qp_solves / tt, nmpc_solves_per_s=B / tt,
            converged_flag_per_sqp_iteration=nconv, problems_converged_at_some_iteration=int(ever.sum()),
            ipm_status_counts=statuses.tolist(),
            ipm_iteration_histogram={str(i): int(v) for i, v in enumerate(hist) if v}, settings=S4,
            note="every SQP iteration = K1+K2+K3+K4 on the device; host reads back 4096 convergence flags + stats per iteration")
        c.close()

# the same workload (spread 0.25) through the DEVICE-side loop (srbd_sqp_solve): no read-back per iteration, converged
# problems are frozen (K3 skips them), the remaining launches return at once when every problem has converged
w = pkg.workload.srbd_batch(B, N=N, contact_mode="gait", spread=0.25)
c = ctx_for(B, N, S4)
c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); c.sqp_solve(HARD, 15); c.sync()
c.reset_sqp_state(); c.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"]); c.sync()
t0 = time.perf_counter(); sqp_it = c.sqp_solve(HARD, 15); tt = time.perf_counter() - t0
conv = c.download_sqp_state()[1]
out["config4_b4096_n100_full_sqp"]["spread_0.25_device_loop"] = dict(
    seconds=tt, nmpc_solves_per_s=B / tt, qp_solves=int(sqp_it.sum()), qp_solves_per_s=float(sqp_it.sum()) / tt,
    problems_converged=int((conv != 0).sum()), sqp_iterations_mean=float(sqp_it.mean()),
    sqp_iteration_histogram={str(i): int(v) for i, v in enumerate(np.bincount(sqp_it, minlength=16)) if v},
    note="srbd_sqp_solve: one call, one read-back; a problem leaves the loop at its first 'nmpc solve success' like the reference")
c.close()

# ---- config 5: one QP, N=50, latency host->host ------------------------------------------------------------------------
N = 50
S5 = dict(S8, iter_max=50)
w = pkg.workload.srbd_batch(1, N=N, contact_mode="stance")
c = ctx_for(1, N, S5)
sx, su = np.zeros((1, N + 1, 12)), np.zeros((1, N, 12)); it = np.zeros(1, dtype=np.int32); stt = np.zeros(1, dtype=np.int32)
for _ in range(10): c.solve_host_graph(HARD, w["x"], w["u"], w["xref"], w["x0"], w["contact"], sx, su, it, stt)
ts = []
for _ in range(1000):
    t0 = time.perf_counter(); c.solve_host_graph(HARD, w["x"], w["u"], w["xref"], w["x0"], w["contact"], sx, su, it, stt); ts.append(time.perf_counter() - t0)
tcs = []
for _ in range(30):
    t0 = time.perf_counter(); o = orc.pipeline(orc.model_params(N), orc.ipm_args(**S5), N, HARD, w["x"], w["u"], w["xref"], w["x0"], w["contact"], threads=1, duals=False); tcs.append(time.perf_counter() - t0)
ts, tcs = np.array(ts) * 1e6, np.array(tcs) * 1e6
out["config5_single_qp_n50_latency"] = dict(gpu_p50_us=float(np.percentile(ts, 50)), gpu_p99_us=float(np.percentile(ts, 99)), ipm_iterations=int(it[0]), status=int(stt[0]),
    cpu_oracle_1core_p50_us=float(np.percentile(tcs, 50)), cpu_oracle_1core_p99_us=float(np.percentile(tcs, 99)), oracle_iterations=int(o["iter"][0]),
    settings=S5,
    note="K1+K2+K3 for one QP on one warp incl. H2D/D2H through srbd_solve_host_graph (one CUDA graph launch per call); "
         "the sequential Riccati chain leaves the GPU no parallelism here")
c.close()
print(json.dumps(out))
