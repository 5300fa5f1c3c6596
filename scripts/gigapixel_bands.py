"""BASELINE config 5 (scaled by --size): SaveImage d=1.0 e=3 of one huge frame, then decode it as contiguous bands of AC-group rows, one
band per rank (no collective on the pixel path; every rank reconstructs one extra group row per side, the host stitches by row
offset). Single process: one GPU encodes the frame, the bands are decoded one after the other on cuda:0 and checked against the full-frame
decode. Under torchrun (WORLD_SIZE ranks, one GPU each): SHARDED ENCODE — every rank builds and encodes only its own band of whole LF-group
rows (JxlB200BandEncoder*; two host reductions: 2 flag bits and the token histograms), rank 0 assembles the file and broadcasts it — then
every rank decodes its own band and rank 0 reports the max-over-ranks times.
Usage: python scripts/gigapixel_bands.py [--size 16384] [--bands 8] [--check] [--effort 3]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=16384); ap.add_argument("--bands", type=int, default=8); ap.add_argument("--check", action="store_true")
ap.add_argument("--effort", type=int, default=3)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
import torch
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
import pkgload
from synth import synthetic_image
P = pkgload.load()
n = args.size
tile = synthetic_image(4096, 4096, seed=0)
reps = -(-n // 4096)


def frame_rows(a, b):
    """BGRA rows [a, b) of the synthetic n x n frame (the 4096 x 4096 tile repeated)."""
    rows = np.concatenate([tile[(y0 % 4096):min(4096, (y0 % 4096) + (b - y0))] for y0 in _starts(a, b)], axis=0)
    rows = np.tile(rows, (1, reps, 1))[:, :n]
    out = np.empty((b - a, n, 4), np.uint8); out[..., 0], out[..., 1], out[..., 2], out[..., 3] = rows[..., 2], rows[..., 1], rows[..., 0], 255
    return out


def _starts(a, b):
    y = a
    while y < b:
        yield y
        y = min(b, (y // 4096 + 1) * 4096)


opts = P.EncoderOptions(quality=90, effort=args.effort)
if world == 1:
    bgra = frame_rows(0, n)
    t = time.time(); data = P.encode_to_memory(bgra, opts); t_enc = time.time() - t
    del bgra
    how = "one GPU"
else:
    y0, rows_mine = P.encode_band_partition(n, world)[rank]
    first, ht, hb = P.band_rows_with_halo(n, y0, rows_mine)
    band = frame_rows(first, y0 + rows_mine + hb) if rows_mine else None
    dist.barrier()
    t = time.time(); data, enc_ms = P.encode_band_distributed(band, n, n, y0, rows_mine, opts, dist, device=local); t_enc = time.time() - t
    del band
    tt = torch.tensor([t_enc, enc_ms / 1e3], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t_enc, dev_s = float(tt[0].item()), float(tt[1].item())
    box = [data]; dist.broadcast_object_list(box, src=0); data = box[0]
    how = "sharded over %d GPUs (bands of LF-group rows; %.2f s of it on the device, max over ranks; exchanged between ranks: 2 flag bits + %.1f MB of histograms, then the sections to rank 0)" % (world, dev_s, 8 * (7425 + 64) * 128 / 1e6)
mp = n * n / 1e6
w, h, gdim, rows = P.band_layout(data)
if rank == 0:
    print("encode on %s: %.0f MP in %.2f s = %.0f MP/s, %.1f MB, %.3f bpp; %d group rows of %d px" % (how, mp, t_enc, mp / t_enc, len(data) / 1e6, len(data) * 8 / n / n, rows, gdim), flush=True)
nb = world if world > 1 else args.bands
parts = P.band_partition(rows, nb)
mine = [parts[rank]] if world > 1 else parts
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
t = time.time(); outs = [P.decode_band(data, a, b, device=local) for a, b in mine if b > a]; torch.cuda.synchronize(); dt = time.time() - t
if dist is not None:
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
free_b, total_b = torch.cuda.mem_get_info()
used = torch.tensor([float(total_b - free_b)], dtype=torch.float64, device="cuda")
if dist is not None:
    dist.all_reduce(used, op=dist.ReduceOp.MAX)
if rank == 0:
    print("device memory in use per rank after the band decode (buffers + caching pools, max over ranks): %.2f GB" % (float(used.item()) / 1e9), flush=True)
if rank == 0:
    halo = sum((min(b + 1, rows) - max(a - 1, 0)) for a, b in parts if b > a) / rows - 1.0
    print("decode as %d bands (%s): %.2f s = %.0f MP/s; redundant reconstruction %.1f %%; bytes exchanged between ranks: 0" % (nb, "one per rank" if world > 1 else "sequential on one GPU", dt, mp / dt, 100 * halo), flush=True)
if args.check and world == 1:
    t = time.time(); full = P.decode_band(data, 0, rows); t_full = time.time() - t
    print("full-frame decode: %.2f s = %.0f MP/s; bands == full frame: %s" % (t_full, mp / t_full, bool(np.array_equal(np.concatenate(outs, axis=0), full))), flush=True)
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
