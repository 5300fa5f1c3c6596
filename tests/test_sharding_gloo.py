"""world_size-2 gloo test of the multi-GPU host logic: file i -> rank i mod N, no data-path collective, max-over-ranks timing."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import pkgload
    P = pkgload.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = P.shard_indices(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)   # per-rank step time; the job time is the max
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    units = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(units)
    if rank == 0:
        out.put((gathered, float(t.item()), float(units.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_covers_all_files_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax, units = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(i for s in gathered for i in s) == list(range(11))
    assert gathered[0] == [0, 2, 4, 6, 8, 10] and gathered[1] == [1, 3, 5, 7, 9]
    assert tmax == 15.0 and units == 11.0
