"""N > 1 host logic on CPU (gloo, world size 2): jobs are sharded across ranks with no data-path
collective; the only collectives are the timing barrier and the max-over-ranks reduction bench.py uses.
Each rank 'commits' its own jobs with the CPU oracle standing in for the device (test infrastructure), and
rank 0 checks that the union of the shards equals the unsharded result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_jobs(n_jobs, rank, world):
    """round-robin job -> rank map (what one-worker-per-GPU popping from a shared queue converges to)"""
    return [j for j in range(n_jobs) if j % world == rank]


def _worker(rank, world, port, n_jobs, out):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import p2oracle as O
    from util import rand_felts

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_jobs(n_jobs, rank, world)
    caps = torch.zeros((n_jobs, 4, 4), dtype=torch.int64)
    dist.barrier()
    for j in mine:
        cols = [rand_felts(1000 * j + c, 1 << 5) for c in range(6)]
        cap = O.batch_from_values(cols, 3, 2, want_leaves=False, want_digests=False)["cap"]
        caps[j] = torch.from_numpy(cap.view(np.int64))
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # bench.py: max over ranks
    dist.all_reduce(caps, op=dist.ReduceOp.SUM)   # test-only gather of the per-job results
    if rank == 0:
        out.put((caps.numpy().view(np.uint64), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_jobs_shard_across_two_ranks():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import p2oracle as O
    from util import rand_felts

    n_jobs, world = 5, 2
    assert sorted(shard_jobs(n_jobs, 0, world) + shard_jobs(n_jobs, 1, world)) == list(range(n_jobs))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_jobs, q)) for r in range(world)]
    for p in procs:
        p.start()
    caps, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 3.0  # rank 0 got jobs 0,2,4
    for j in range(n_jobs):
        cols = [rand_felts(1000 * j + c, 1 << 5) for c in range(6)]
        ref = O.batch_from_values(cols, 3, 2, want_leaves=False, want_digests=False)["cap"]
        assert (caps[j] == ref).all()
