"""Multi-GPU partitioning of an ensemble: trajectories are independent, so the
index range [0, N) is cut into contiguous shards, one per GPU / rank, with NO data-path
collective; results are gathered on the host.  Inside one process libb200ens does this itself
(b200ens_opts.device_mask, one host thread + two streams per device).  This module is the
one-process-per-GPU flavour used under torchrun (bench.py --gpus N): every rank solves its
shard with `traj_offset = lo` (so Philox counters and parameter rows are those of the global
ensemble) and rank 0 optionally gathers the host results.
"""
import numpy as np


def shard_range(N, rank, world):
    """Contiguous range [lo, hi) of trajectory indices owned by `rank` (same split as b200ens_solve)."""
    return N * rank // world, N * (rank + 1) // world


def solve_sharded(local_solve, N, u0, p, rank, world, gather=True, group=None):
    """Run `local_solve(u0_shard, p_shard, traj_offset) -> (out, retcode, stats)` on this rank's shard and
    gather the host arrays on rank 0 (torch.distributed object gather over the CPU/gloo group; the data path
    itself has no collective).  Returns (out, retcode, stats) on rank 0, else the local triple."""
    lo, hi = shard_range(N, rank, world)
    out, rc, st = local_solve(u0[lo:hi], p[lo:hi], lo)
    if world == 1 or not gather:
        return out, rc, st
    import torch.distributed as dist

    parts = [None] * world if rank == 0 else None
    dist.gather_object((lo, hi, out, rc, st), parts, dst=0, group=group)
    if rank != 0:
        return out, rc, st
    parts.sort(key=lambda x: x[0])
    assert parts[0][0] == 0 and parts[-1][1] == N and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    cat = lambda k: np.concatenate([x[k] for x in parts], axis=0)
    return cat(2), cat(3), (cat(4) if parts[0][4] is not None else None)


def allreduce_summary(summary, group=None):
    """Merge per-rank EnsembleSummary partial sums across ranks: the ONE collective of this back-end (SURVEY 8(f)
    item 2) -- a few hundred doubles, all-reduced with torch.distributed (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return summary
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    buf = torch.from_numpy(np.concatenate([summary.sum.ravel(), summary.sumsq.ravel(), [float(summary.num_monte)]])).to(dev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    h = buf.cpu().numpy()
    k = summary.sum.size
    summary.sum = h[:k].reshape(summary.sum.shape)
    summary.sumsq = h[k:2 * k].reshape(summary.sumsq.shape)
    summary.num_monte = int(round(h[2 * k]))
    return summary
