"""CPU tests (-m "not gpu") of the N>1 host logic with world_size 2 over gloo: the shards of a batch are
disjoint, cover it, reproduce the single-rank inputs bit for bit, and the gathered statistics block equals the
single-rank one (solves, iteration sum and histogram, status counts, max residuals)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, N = 48, 8
SETTINGS = dict(iter_max=30, alpha_min=1e-8, mu0=1e2, tol_stat=1e-8, tol_eq=1e-8, tol_ineq=1e-8, tol_comp=1e-8,
                reg_prim=1e-12, warm_start=0, pred_corr=1, ric_alg=0, split_step=1)


def _stats_from(out):
    it, st = out["iter"], out["status"]
    return dict(solves=len(it), iter_sum=int(it.sum()), status_count=[int((st == s).sum()) for s in range(5)],
                iter_hist=[int(v) for v in np.bincount(np.minimum(it, 63), minlength=64)],
                res_max=[float(v) for v in out["res_max"].max(axis=0)])


def _solve_shard(start, count):
    """Stand-in for the per-rank GPU solve: the oracle on the rank's shard (the GPU path is checked against the
    oracle in tests/test_gpu_parity.py)."""
    sys.path.insert(0, ROOT)
    import srbd_pkg
    pkg = srbd_pkg.load()
    from oracle import oracle as orc
    w = pkg.workload.srbd_batch(count, N=N, contact_mode="gait", start=start)
    out = orc.pipeline(orc.model_params(N), orc.ipm_args(**SETTINGS), N, 1, w["x"], w["u"], w["xref"], w["x0"],
                       w["contact"], threads=2)
    return pkg, w, out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sys.path.insert(0, ROOT)
    import srbd_pkg
    pkg = srbd_pkg.load()
    start, count = pkg.sharding.shard_range(B, rank, world)
    _, w, out = _solve_shard(start, count)
    total, per_rank = pkg.sharding.gather_batch_stats(_stats_from(out))
    tmax = pkg.sharding.max_over_ranks(10.0 + rank)
    q.put((rank, start, count, total, per_rank, tmax, w["x0"].tobytes(), out["iter"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pkg, w, out = _solve_shard(0, B)
    single = _stats_from(out)
    (r0, s0, c0, tot0, pr0, tmax0, x00, it0), (r1, s1, c1, tot1, pr1, tmax1, x01, it1) = res
    assert (s0, c0, s1, c1) == (0, 24, 24, 24)
    assert x00 + x01 == w["x0"].tobytes()               # shards reproduce the single-rank inputs bit for bit
    assert it0 + it1 == out["iter"].tolist()            # and therefore the same per-QP iteration counts
    assert tot0 == tot1                                 # every rank holds the same gathered block
    for k in ("solves", "iter_sum", "status_count", "iter_hist"):
        assert tot0[k] == single[k], k
    assert np.allclose(tot0["res_max"], single["res_max"], rtol=0, atol=0)
    assert pr0[0]["solves"] == 24 and pr0[1]["solves"] == 24
    assert tmax0 == tmax1 == 11.0


@pytest.mark.parametrize("total,world", [(65536, 8), (10, 4), (7, 8), (1, 1)])
def test_shard_range(total, world):
    sys.path.insert(0, ROOT)
    import srbd_pkg
    pkg = srbd_pkg.load()
    spans = [pkg.sharding.shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == total
    for (s, c), (s2, _) in zip(spans, spans[1:]):
        assert s + c == s2
    assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
