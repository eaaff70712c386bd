"""N>1 host logic on CPU: world_size-2 gloo run of the one-process-per-GPU sharding path
(differentialequations.jl_b200/sharding.py) with the oracle standing in for the per-rank device solve.
Checks: shards are disjoint and cover [0,N), traj_offset makes Philox noise identical to the single-process
run, the host gather reassembles the ensemble in order."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import torch.distributed as dist
    import oracle_py
    import b200ens
    from b200ens import sharding, workloads as W

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    u0, p = W.gbm_params(N)

    def local(u0s, ps, off):
        out, rc, st = oracle_py.solve("gbm", "EM", u0s, ps, (0.0, 1.0), [0.5, 1.0], 1 / 32, seed=99, adaptive=False,
                                      traj_offset=off, nthreads=1)
        return out, rc, st

    out, rc, st = sharding.solve_sharded(local, N, u0, p, rank, world)
    # the one collective of the back-end: all-reduce of per-rank ensemble moments (gloo here, NCCL on GPUs)
    lo, hi = sharding.shard_range(N, rank, world)
    mine = out if world == 1 or rank != 0 else out[lo:hi]
    part = b200ens.EnsembleSummary(np.array([0.5, 1.0]), mine.sum(axis=0), (mine * mine).sum(axis=0), mine.shape[0],
                                   np.ones(mine.shape[0], dtype=np.int32), 0.0, {})
    summ = sharding.allreduce_summary(part)
    if rank == 0:
        q.put((out, rc, st, summ.u, summ.v, summ.num_monte))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    sys.path.insert(0, ROOT)
    import b200ens
    from b200ens.sharding import shard_range

    for N in (1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [shard_range(N, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == N and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_world2_gloo_matches_single_process(oracle):
    from b200ens import workloads as W

    N, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    out, rc, st, mean, var, cnt = q.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    u0, p = W.gbm_params(N)
    ref, rc1, _ = oracle.solve("gbm", "EM", u0, p, (0.0, 1.0), [0.5, 1.0], 1 / 32, seed=99, adaptive=False)
    assert out.shape == (N, 2, 1) and np.array_equal(out, ref) and np.array_equal(rc, rc1)
    assert cnt == N and np.allclose(mean, ref.mean(axis=0), rtol=1e-12) and np.allclose(var, ref.var(axis=0, ddof=1), rtol=1e-9)
