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


def _band_worker(rank, world, port, data, out):
    """Gigapixel-style sharding of ONE frame: every rank parses the layout from the same file bytes (host only, no GPU
    needed), takes its contiguous group-row range, and the host stitches the bands by row offset. No collective carries pixel data."""
    sys.path.insert(0, ROOT)
    import pkgload
    P = pkgload.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, gdim, rows = P.band_layout(data)
    begin, end = P.band_partition(rows, world)[rank]
    mine = (begin * gdim, min(end * gdim, h)) if end > begin else (0, 0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, begin, end, mine))
    if rank == 0:
        out.put(((w, h, gdim, rows), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_sharding_of_one_frame():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    data = O.encode(O.synthetic_image(300, 1100, seed=4), effort=3)      # 2 x 5 groups of 256 px
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, data, q)) for r in range(2)]
    for p in procs:
        p.start()
    layout, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert layout == (300, 1100, 256, 5)
    assert gathered == [(0, 0, 3, (0, 768)), (1, 3, 5, (768, 1100))]       # contiguous, disjoint, covering every image row


def _enc_worker(rank, world, port, out):
    """Host logic of the sharded encoder (P.encode_band_distributed) with the GPU session replaced by a recorder: the partition, the OR of
    the band flags, the SUM of the histograms and the gather of the section blobs are what is under test."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import pkgload
    P = pkgload.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    width, height = 64, 5000
    seen = {}

    class FakeBand:
        def __init__(self, rows, frame_height, first_row, halo_top, halo_bottom, options, device=-1):
            seen["args"] = (rows.shape[0], frame_height, first_row, halo_top, halo_bottom)
            self.flags = 1 if first_row == 0 else 2           # rank 0 saw colour, rank 1 saw transparency
            self.device_ms = 1.0

        def tokenize(self, frame_flags):
            seen["flags"] = frame_flags
            return np.arange(16, dtype=np.uint64) * np.uint64(seen["args"][2] // 2048 + 1)

        def finish(self, total):
            seen["total"] = total.copy()
            return b"band@%d" % seen["args"][2]

        def close(self):
            pass

    P.BandEncoder = FakeBand
    P.assemble_bands = lambda w, h, o, flags, total, blobs, metadata=None: (w, h, flags, [int(v) for v in total], blobs)
    y0, rows = P.encode_band_partition(height, world)[rank]
    first, ht, hb = P.band_rows_with_halo(height, y0, rows)
    surface = np.zeros((height, width, 4), np.uint8)
    result, ms = P.encode_band_distributed(surface[first:y0 + rows + hb], width, height, y0, rows, P.EncoderOptions(quality=90, effort=7), dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seen["args"], seen["flags"], [int(v) for v in seen["total"]]))
    if rank == 0:
        out.put((result, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_encode_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_enc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    result, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # 5000 rows = 3 LF-group rows: rank 0 takes two (rows 0..4095), rank 1 the last (4096..4999)
    assert gathered[0][0] == (4096 + 8, 5000, 0, 0, 8) and gathered[1][0] == (904 + 8, 5000, 4096, 8, 0)
    want = [i * 1 + i * 3 for i in range(16)]                             # band 0 reports 1 * arange, band 1 (LF-group row 2) 3 * arange
    assert gathered[0][1] == 3 and gathered[1][1] == 3                    # flags OR-ed over the bands
    assert gathered[0][2] == want and gathered[1][2] == want              # histograms summed over the bands
    assert result == (64, 5000, 3, want, [b"band@0", b"band@4096"])       # rank 0 assembles the blobs in rank order
