"""Multi-GPU plumbing: independent QPs shard over ranks, one process per GPU, no data-path
collective; the only exchange is the gather of the per-rank statistics block that K3 reduces on the device
(`srbd_batch_stats`).  Works with the NCCL backend (CUDA tensors over NVLink) and with gloo (CPU tests)."""
import numpy as np

HIST_BINS = 64


def shard_range(total, rank, world):
    """Contiguous block split: rank r owns QPs [start, start+count)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def stats_to_vector(bs):
    """dict from Context.batch_stats() -> (int64 vector [solves, iter_sum, status x5, hist x64], float64 res_max[4])."""
    iv = np.array([bs["solves"], bs["iter_sum"]] + list(bs["status_count"]) + list(bs["iter_hist"]), dtype=np.int64)
    return iv, np.array(bs["res_max"], dtype=np.float64)


def vector_to_stats(iv, fv):
    iv = [int(v) for v in iv]
    return dict(solves=iv[0], iter_sum=iv[1], status_count=iv[2:7], iter_hist=iv[7:7 + HIST_BINS],
                res_max=[float(v) for v in fv])


def gather_batch_stats(bs, device=None):
    """All-gather the statistics blocks of every rank and reduce them (sum of counters, max of residuals).
    Returns (global stats dict, list of per-rank dicts).  Single-process: returns the input."""
    import torch
    import torch.distributed as dist
    iv, fv = stats_to_vector(bs)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vector_to_stats(iv, fv), [vector_to_stats(iv, fv)]
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    ti = torch.from_numpy(iv).to(dev)
    tf = torch.from_numpy(fv).to(dev)
    gi = [torch.zeros_like(ti) for _ in range(dist.get_world_size())]
    gf = [torch.zeros_like(tf) for _ in range(dist.get_world_size())]
    dist.all_gather(gi, ti)
    dist.all_gather(gf, tf)
    per_rank = [vector_to_stats(a.cpu().numpy(), b.cpu().numpy()) for a, b in zip(gi, gf)]
    tot_i = torch.stack(gi).sum(0).cpu().numpy()
    tot_f = torch.stack(gf).max(0).values.cpu().numpy()
    return vector_to_stats(tot_i, tot_f), per_rank


def max_over_ranks(value, device=None):
    """Device-timed milliseconds -> max over ranks (the number every multi-GPU figure is quoted on)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
