#!/usr/bin/env python
"""bench.py — SRBD NMPC QP solves/sec (BASELINE.json metric) on N B200s of one node.

One "step" = one pass of the hot path (K1 linearize -> K2 assemble -> K3 OCP-QP IPM solve to tol 1e-8)
over one batch of synthetic SRBD problems: BASELINE config 3 (65536 QPs per GPU, horizon N=20, randomized
contact schedules, hard friction-cone / force-box rows).  Independent QPs shard over ranks with no
data-path collective (weak scaling: 65536 QPs per GPU); NCCL only gathers the per-rank statistics block.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of the reference path on the host cores

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same metric through the C-ABI with pinned HOST buffers (H2D + solve + D2H in the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SETTINGS = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
                reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)
METRIC = "SRBD NMPC QP solves/sec, batch 64k, N=20"
UNIT = "solves/s"
HORIZON = 20


def stage_flops(N, nx=12, nu=12, ng=24):
    """Algorithmic flops of SURVEY.md §8d: returns (F_it with c=2 KKT vector solves, F_res) per QP."""
    f_it, f_res = 0.0, 0.0
    for k in range(N + 1):
        nuk, nxk = (nu if k < N else 0), (nx if k > 0 else 0)
        n, nxn, ngk = nuk + nxk, (nx if k < N else 0), (ng if k < N else 0)
        fact = nxn * nxn * (n + 1) + n * (n + 1) * (nxn + ngk) + n ** 3 / 3.0 + n * n + n * ngk
        vec = 2.0 * (n * n + 2 * n * nxn + 2 * nxn * nxn) + 2.0 * n * ngk
        res = 2.0 * n * n + 4.0 * n * nxn + 4.0 * n * ngk
        f_it += fact + 2.0 * vec + res
        f_res += res
    return f_it, f_res


def bench_config(world, B, contact):
    """The `config` object of the JSON line, shared by the B200 arm and the --impl reference arm (same workload)."""
    return {"workload": "BASELINE config 3: 65536 SRBD QPs per GPU, N=20, randomized contact schedules, "
                        "hard friction-cone/force-box rows (HARD_INEQ), full IPM to tol_stat=tol_eq=tol_ineq=tol_comp=1e-8, "
                        "ric_alg=0",
            "qps_per_gpu": B, "global_batch": world * B, "horizon": HORIZON, "contact": contact,
            "l2": "inputs larger than L2 (packed QP data per step >> 126 MB, trajectories 0.4 GB)",
            "parallelism": f"dp{world} (independent QPs, no data-path collective)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            try:
                pw.append(float(r[3]))
            except Exception:
                pass
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_min": float(min(sm)), "power_w": float(np.median(pw)) if pw else None}


def cpu_reference_run(pkg, batch, threads=0, seed_start=0, contact="gait"):
    """Times the CPU oracle (restatement of the reference path) on `batch` QPs.  Returns (solves/s, cores, out)."""
    from oracle import oracle as orc
    w = pkg.workload.srbd_batch(batch, N=HORIZON, contact_mode=contact, start=seed_start)
    m, a = orc.model_params(HORIZON), orc.ipm_args(**SETTINGS)
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would make
    # the reported CPU baseline a single-core number at N > 1
    if threads <= 0:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    nthr = threads
    t0 = time.perf_counter()
    out = orc.pipeline(m, a, HORIZON, pkg.capi.SRBD_HARD_INEQ, w["x"], w["u"], w["xref"], w["x0"], w["contact"],
                       threads=nthr, duals=False)
    dt = time.perf_counter() - t0
    return batch / dt, nthr, out, dt


def cpu_baseline_record(value, cores, sample, **extra):
    """cpu_baseline object: the oracle PORT (plain-C restatement with naive loops), not the reference's HPIPM/BLASFEO
    (not buildable offline: no Eigen / yaml-cpp / HPIPM / BLASFEO sources).  BLASFEO's hand-written AVX2 / AVX-512
    micro-kernels are several times faster than these triple loops on the same cores: read the GPU/CPU ratio with that
    in mind (the roofline fraction, not this ratio, measures kernel quality)."""
    rec = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
           "solves_per_s_per_core": value / max(cores, 1),
           "build": "gcc -O3 -mavx2 -mfma -fopenmp (oracle/Makefile), plain triple loops, one QP per OpenMP task",
           "note": "CPU oracle port; the reference's HPIPM/BLASFEO/Eigen cannot be built offline and would be several "
                   "times faster per core (hand-written BLASFEO micro-kernels)"}
    rec.update(extra)
    return rec


def run_reference(args, pkg, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself cannot be built
    offline (Eigen / yaml-cpp / HPIPM / BLASFEO absent), so this is the oracle port (kind = "port")."""
    if rank != 0:
        return
    sample = args.cpu_sample
    for _ in range(args.warmup):
        cpu_reference_run(pkg, min(sample, 512))
    times, cores = [], 1
    for s in range(args.steps):
        v, cores, out, dt = cpu_reference_run(pkg, sample, seed_start=s * sample)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            # the SAME workload as the B200 arm (config 3); each timed step is a bounded sample of it (cpu_baseline.sample)
            "config": bench_config(world, args.batch, args.contact),
            "cpu_baseline": cpu_baseline_record(val, cores, f"{sample} consecutive QPs of that workload per step x "
                                                f"{args.steps} steps (a different slice each step), OpenMP over QPs"),
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json(line)


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything else any library prints there (the NCCL version banner comes
    from C code, whatever NCCL_DEBUG says) is sent to stderr; emit_json() writes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def latency_leg(pkg, reps=300):
    """The second half of BASELINE.json's metric ("... ; p50 solve latency"), config 5: ONE SRBD QP, N=50, all stance,
    HARD_INEQ, all four tolerances 1e-8, solved host memory -> host memory, 10 warm-ups, then `reps` back-to-back calls
    timed on the host clock (the calls are synchronous).  Two forms of the same pipeline (K1 + K2 + K3 + copies):
    srbd_solve_host_graph (pinned staging in the context, ONE CUDA graph launch per call) and the plain srbd_solve_host
    (five copies up, five launches, four copies down)."""
    from srbd_nmpc_solver_b200.binding import make_dims
    N = 50
    w = pkg.workload.srbd_batch(1, N=N, contact_mode="stance")
    c = pkg.Context(1, make_dims(N=N))
    c.set_model(pkg.default_model_params(N))
    c.set_ipm_args(pkg.default_ipm_args(**dict(SETTINGS, iter_max=50)))
    sx, su = np.zeros((1, N + 1, 12)), np.zeros((1, N, 12))
    it, stt = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
    mode = pkg.capi.SRBD_HARD_INEQ
    out = {}
    for name, fn in (("graph", c.solve_host_graph), ("plain", c.solve_host)):
        for _ in range(10):
            fn(mode, w["x"], w["u"], w["xref"], w["x0"], w["contact"], sx, su, it, stt)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn(mode, w["x"], w["u"], w["xref"], w["x0"], w["contact"], sx, su, it, stt)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e6
        out[name] = (float(np.percentile(ts, 50)), float(np.percentile(ts, 99)), sx.copy())
    c.close()
    assert np.array_equal(out["graph"][2], out["plain"][2]), "graph replay and plain call disagree"
    # the CPU oracle on the same problem, one core (the reference calls HPIPM once per control step, NMPC_solver.cpp:316-330)
    from oracle import oracle as orc
    m, a = orc.model_params(N), orc.ipm_args(**dict(SETTINGS, iter_max=50))
    tc = []
    for _ in range(5):
        t0 = time.perf_counter()
        o = orc.pipeline(m, a, N, mode, w["x"], w["u"], w["xref"], w["x0"], w["contact"], threads=1, duals=False)
        tc.append(time.perf_counter() - t0)
    return {"p50_us": out["graph"][0], "p99_us": out["graph"][1], "calls": reps,
            "p50_us_plain_call": out["plain"][0], "p99_us_plain_call": out["plain"][1],
            "cpu_oracle_port_1core_us": float(np.median(tc) * 1e6), "cpu_iterations": int(o["iter"][0]),
            "ipm_iterations": int(it[0]), "status": int(stt[0]),
            "workload": "BASELINE config 5: one SRBD QP, N=50, all stance, HARD_INEQ, tol_stat=tol_eq=tol_ineq=tol_comp=1e-8, "
                        "host->host through srbd_solve_host_graph on rank 0",
            "note": "one QP = one warp: the solve is a serial Riccati chain (51 stages x 12 iterations x 5 sweeps); the "
                    "GPU path is a throughput design (DESIGN.md section 5, latency)"}


def facade_leg(batch, device):
    """The reference's OWN boundary as the measured path: hpipm::OcpQpIpmSolver::solveBatch on host std::vector<OcpQp>
    (srbd-nmpc-solver_b200/host/tests/bench_facade.cpp, built by __graft_entry__.build()).  The QP-level interface moves
    228 KB per N=20 QP (every A, B, Q, S, R, C, D; the all-zero S and C stay on the host: 161 KB) against 6 KB for the
    NMPC-level calls, so a 65536-QP batch would be 15 GB of host data: measured at `batch` QPs and reported per second.
    solveBatch pipelines the batch in chunks of 1024 QPs over two pooled contexts (hpipm-cpp.hpp: solvePipelined)."""
    exe = os.path.join(ROOT, "srbd-nmpc-solver_b200", "host", "tests", "bench_facade")
    if not os.path.exists(exe):
        return {"unavailable": "bench_facade not built (python -c 'import __graft_entry__ as g; g.build()')"}
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[device]
               if os.environ.get("CUDA_VISIBLE_DEVICES") else str(device))
    try:
        r = subprocess.run([exe, str(batch), "2"], capture_output=True, text=True, timeout=600, env=env)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # the headline line must not depend on this leg
        return {"unavailable": repr(e)[:200]}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="QPs per GPU")
    ap.add_argument("--contact", default="gait", choices=["gait", "stance"])
    ap.add_argument("--cpu-sample", type=int, default=8192, help="QPs per step of the CPU baseline / reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-instance latency leg (BASELINE config 5)")
    ap.add_argument("--no-facade", action="store_true", help="skip the hpipm-cpp facade leg (host/tests/bench_facade)")
    ap.add_argument("--facade-batch", type=int, default=16384)
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import srbd_pkg
    pkg = srbd_pkg.load()

    if args.impl == "reference":
        run_reference(args, pkg, rank, world)
        return

    # NCCL_DEBUG is left as the caller set it: whatever NCCL prints on stdout goes to stderr (claim_stdout above), so
    # the driver can count ranks in the NCCL log and stdout still carries exactly one JSON line
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = args.batch
    w = pkg.workload.srbd_batch(B, N=HORIZON, contact_mode=args.contact, start=rank * B)
    ctx = pkg.Context(B, device=local_rank)
    ctx.set_model(pkg.default_model_params(HORIZON))
    ctx.set_ipm_args(pkg.default_ipm_args(**SETTINGS))
    # Roofline denominator FIRST, on a cool board: a short DMMA burst (srbd_fp64_peak).  Measured after the sustained
    # loop the same probe reads 30-35 TFLOP/s (1000 W cap), which made round 1's `frac` move between runs.
    fp64_peak = ctx.fp64_peak() / 1e12
    ctx.upload_traj(w["x"], w["u"], w["xref"], w["x0"], w["contact"])
    ctx.sync()
    mode = pkg.capi.SRBD_HARD_INEQ

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(c, steps, warmup):
        """W untimed + K timed steps (K1 -> K2 -> K3) of context c with device-resident inputs; CUDA events on the
        context's stream.  Returns (ms over the K steps on this rank, per-step [K1, K2, K3] ms, launches)."""
        stream = torch.cuda.ExternalStream(c.stream, device=local_rank)

        def step(evs=None):
            if evs:
                evs[0].record(stream)
            c.linearize()
            if evs:
                evs[1].record(stream)
            c.assemble(mode)
            if evs:
                evs[2].record(stream)
            c.qp_solve()
            if evs:
                evs[3].record(stream)
        for _ in range(warmup):
            step()
        c.sync()
        barrier()
        l0 = c.launch_count
        evs = [[ev() for _ in range(4)] for _ in range(steps)]
        for s_ in range(steps):
            step(evs[s_])
        barrier()
        return (evs[0][0].elapsed_time(evs[-1][3]),
                np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in evs]), c.launch_count - l0)

    sampler = ClockSampler(local_rank)   # every rank samples ITS GPU during the timed region
    sampler.start()
    total_ms, k_ms, launches = timed_steps(ctx, args.steps, args.warmup)
    clocks_rank = sampler.stop()
    clocks = clocks_rank if rank == 0 else None
    total_ms_max = pkg.sharding.max_over_ranks(total_ms)  # device time, max over ranks
    ms_per_step = total_ms_max / args.steps
    value = world * B / (ms_per_step * 1e-3)
    # per-rank record: names the limiter when the max-over-ranks time is not rank 0's
    mine = {"rank": rank, "ms_per_step": total_ms / args.steps, "k3_ms": float(k_ms[:, 2].mean()),
            "sm_mhz": clocks_rank.get("sm_mhz"), "sm_mhz_min": clocks_rank.get("sm_mhz_min"),
            "power_w": clocks_rank.get("power_w"), "reasons": clocks_rank.get("reasons")}
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    else:
        per_rank = [mine]
    rank_ms = np.array([r["ms_per_step"] for r in per_rank])

    # strong scaling of BASELINE config 3 AS WRITTEN ("batch 65536 ... sharded over 1/2/4/8"): 65536 QPs in total,
    # 65536 / N per GPU (rank r owns QPs [r B/N, (r+1) B/N) of the same generator)
    strong = None
    if world > 1:
        Bs = B // world
        ws_ = pkg.workload.srbd_batch(Bs, N=HORIZON, contact_mode=args.contact, start=rank * Bs)
        cs = pkg.Context(Bs, device=local_rank)
        cs.set_model(pkg.default_model_params(HORIZON))
        cs.set_ipm_args(pkg.default_ipm_args(**SETTINGS))
        cs.upload_traj(ws_["x"], ws_["u"], ws_["xref"], ws_["x0"], ws_["contact"])
        cs.sync()
        s_ms, _, _ = timed_steps(cs, args.steps, args.warmup)
        s_ms_max = pkg.sharding.max_over_ranks(s_ms) / args.steps
        strong = {"scaling": "strong", "global_batch": Bs * world, "qps_per_gpu": Bs, "ms_per_step": s_ms_max,
                  "value": Bs * world / (s_ms_max * 1e-3), "unit": UNIT,
                  "note": "value(N) / value(1 GPU, 65536 QPs) / N is the strong-scaling efficiency; a 65536 / N shard is "
                          "65536 / N / 1776 resident-QP waves of the persistent K3 grid (tail effects grow with N)"}
        cs.close()

    # per-rank statistics block (fused epilogue of K3), gathered over NCCL
    bs = ctx.batch_stats()
    st = ctx.download_stats()
    stats_all, _ = pkg.sharding.gather_batch_stats(bs)

    # ---- e2e: host buffers through the C-ABI (pinned), H2D + K1 + K2 + K3 + D2H per step ---------------
    def pinned(a):
        tt = torch.from_numpy(a).pin_memory()
        return tt, tt.numpy()
    keep, hb = [], {}
    for k in ("x", "u", "xref", "x0", "contact"):
        tt, hb[k] = pinned(w[k])
        keep.append(tt)
    # Streaming use of the C-ABI: TWO contexts (two CUDA streams) take the steps alternately through
    # srbd_solve_host_async / srbd_wait, so the H2D / D2H copies of one step overlap the kernels of the other.  Every
    # step still moves its own inputs from pinned host memory and its own results back to pinned host memory.
    ctx2 = pkg.Context(B, device=local_rank)
    ctx2.set_model(pkg.default_model_params(HORIZON))
    ctx2.set_ipm_args(pkg.default_ipm_args(**SETTINGS))
    lanes = []
    for c in (ctx, ctx2):
        outs = [pinned(np.zeros((B, HORIZON + 1, 12))), pinned(np.zeros((B, HORIZON, 12))),
                pinned(np.zeros(B, dtype=np.int32)), pinned(np.zeros(B, dtype=np.int32))]
        keep.extend(o[0] for o in outs)
        lanes.append((c, [o[1] for o in outs]))

    def submit(lane):
        c, (o_sx, o_su, o_it, o_st) = lane
        c.solve_host_async(mode, hb["x"], hb["u"], hb["xref"], hb["x0"], hb["contact"], o_sx, o_su, o_it, o_st)

    for lane in lanes:  # warm-up (first-use allocations)
        submit(lane)
        lane[0].wait()
    barrier()
    t0 = time.perf_counter()
    # as many end-to-end steps as device-timed ones (>= 2 x --e2e-steps): the first H2D and the last D2H of the run cannot
    # overlap anything (12 ms of pipeline fill / drain), which six steps would charge at 2 % each
    n_e2e = max(2 * args.e2e_steps, 2 * ((args.steps + 1) // 2))
    submit(lanes[0])
    for i in range(1, n_e2e):
        submit(lanes[i & 1])          # step i is enqueued ...
        lanes[(i - 1) & 1][0].wait()  # ... while step i-1 finishes: its results are on the host now
    lanes[(n_e2e - 1) & 1][0].wait()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    h_sx, h_su, h_it, h_st = lanes[0][1]
    h_it2, h_st2 = lanes[1][1][2], lanes[1][1][3]
    e2e_ms_max = pkg.sharding.max_over_ranks(e2e_ms)
    e2e_value = world * B / (e2e_ms_max * 1e-3)
    h2d = sum(int(hb[k].nbytes) for k in hb)
    d2h = int(h_sx.nbytes + h_su.nbytes + h_it.nbytes + h_st.nbytes)
    # both paths run the same kernels on the same inputs: results must be identical.  (Not every QP has to converge:
    # at tol 1e-8 the IPM sits on its rounding floor for roughly 1 QP in 30000 -- in the CPU oracle as well, on other
    # QPs, see DESIGN.md section 2 -- the status counts are reported below.)
    assert np.array_equal(h_st, st["status"]) and np.array_equal(h_it, st["iter"]), \
        "e2e path disagrees with the device-resident path"
    assert np.array_equal(h_st2, st["status"]) and np.array_equal(h_it2, st["iter"]), \
        "second e2e context disagrees with the device-resident path"

    if rank == 0:
        # roofline of the dominant kernel (K3): algorithmic FP64 flops with the ACTUAL iteration counts
        f_it, f_res = stage_flops(HORIZON)
        flops_launch = float(st["iter"].sum()) * f_it + B * f_res
        k3_ms = float(k_ms[:, 2].mean())
        achieved = flops_launch / (k3_ms * 1e-3) / 1e12
        peak = fp64_peak
        traffic, traffic_src, pipes_ncu = None, None, None
        tpath = os.path.join(ROOT, "profiles", "k3_traffic.json")
        if os.path.exists(tpath):  # dram bytes of K3 from the committed `ncu --set full` capture, scaled per QP
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["dram_bytes_per_qp"] * B, tj["source"]
            pipes_ncu = tj.get("pipes_ncu")
        # "tensor": K3 runs on the FP64 tensor cores (DMMA); their peak IS the FP64 pipe's (DMMA and DFMA share one
        # datapath on B200: scripts/microbench/fp64_pipes.cu), measured live -- not the bf16 figure of MEASURED_PEAKS.json
        roofline = {"bound": "tensor", "pipe": "fp64 (DMMA m8n8k4 + DFMA, one shared datapath)",
                    "kernel": "ipm_srbd_kernel (K3)", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak if peak > 0 else None, "traffic": traffic,
                    "traffic_source": traffic_src, "pipes_ncu": pipes_ncu,
                    "peak_source": "measured live BEFORE the timed loop: DMMA-saturating burst srbd_fp64_peak() "
                                   "(MEASURED_PEAKS.json carries no FP64 figure; DMMA and DFMA share the pipe)",
                    "flops_per_launch": flops_launch, "kernel_ms": k3_ms,
                    "kernel_share_of_step": k3_ms / float(k_ms.sum(1).mean()),
                    "k1_ms": float(k_ms[:, 0].mean()), "k2_ms": float(k_ms[:, 1].mean())}
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
        # HBM view of K1/K2 (write-bound kernels): dense packed records written per step
        k1_bytes = B * HORIZON * (72 + 12) * 8 + B * (HORIZON * 2 + 1) * 12 * 8   # dyn records + defect written (the dense BAbt records are lazy); x, u, x0 read
        k2_bytes = B * (HORIZON + 1) * 192 * 8  # the compact stage records only (dense RSQrq / DCt / d are materialised lazily)
        roofline["k1_hbm_frac"] = k1_bytes / (roofline["k1_ms"] * 1e-3) / 1e9 / hbm_peak
        roofline["k2_hbm_frac"] = k2_bytes / (roofline["k2_ms"] * 1e-3) / 1e9 / hbm_peak
        roofline["hbm_peak_gbs"] = hbm_peak
        # HBM view of K3: its measured DRAM traffic (ncu capture above) over its live kernel time
        if traffic:
            roofline["k3_hbm_gbs"] = traffic / (k3_ms * 1e-3) / 1e9
            roofline["k3_hbm_frac"] = roofline["k3_hbm_gbs"] / hbm_peak
        cpu_baseline = None
        if not args.no_cpu_baseline:
            v, cores, out, dt = cpu_reference_run(pkg, args.cpu_sample)
            same = bool(np.array_equal(out["iter"], st["iter"][:args.cpu_sample])) if args.cpu_sample <= B else None
            cpu_baseline = cpu_baseline_record(v, cores, f"first {args.cpu_sample} QPs of rank 0's shard, one pass "
                                               f"({dt:.1f} s), OpenMP over QPs", iteration_counts_equal_gpu=same)
        latency = None if args.no_latency else latency_leg(pkg)
        facade = None if args.no_facade else facade_leg(args.facade_batch, local_rank)
        it = st["iter"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": bench_config(world, B, args.contact),
                "clocks": clocks, "gpu_launches": int(launches),
                "per_rank_ms": {"min": float(rank_ms.min()), "median": float(np.median(rank_ms)),
                                "max": float(rank_ms.max()), "ranks": per_rank},
                "strong": strong,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms_max, "steps": n_e2e,
                        "how": "two contexts / CUDA streams take the steps alternately (srbd_solve_host_async + srbd_wait): "
                               "copies of one step overlap the kernels of the other"},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "latency": latency, "facade_e2e": facade,
                "ipm": {"iter_mean": float(it.mean()), "iter_min": int(it.min()), "iter_max": int(it.max()),
                        "status_counts_all_ranks": stats_all["status_count"], "solves_all_ranks": stats_all["solves"],
                        "iter_sum_all_ranks": stats_all["iter_sum"], "res_max_all_ranks": stats_all["res_max"]}}
        emit_json(line)
    ctx2.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
