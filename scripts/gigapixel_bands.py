"""BASELINE config 5 (scaled by --size): SaveImage d=1.0 e=3 of one huge frame, then decode it as contiguous bands of AC-group rows, one
band per rank (no collective; every rank parses the same file, reconstructs one extra group row per side, and the host
stitches by row offset). Single process: the bands are decoded one after the other on cuda:0 and checked against the full-frame decode.
Under torchrun (WORLD_SIZE ranks, one GPU each) every rank decodes its own band and rank 0 reports the max-over-ranks time.
Usage: python scripts/gigapixel_bands.py [--size 16384] [--bands 8] [--check]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=16384); ap.add_argument("--bands", type=int, default=8); ap.add_argument("--check", action="store_true")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
import torch
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pkgload
from synth import synthetic_image
P = pkgload.load()
n = args.size
tile = synthetic_image(4096, 4096, seed=0)
reps = -(-n // 4096)
img = np.tile(tile, (reps, reps, 1))[:n, :n]
bgra = np.empty((n, n, 4), np.uint8); bgra[..., 0], bgra[..., 1], bgra[..., 2], bgra[..., 3] = img[..., 2], img[..., 1], img[..., 0], 255
del img
t = time.time(); data = P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=3)); t_enc = time.time() - t
del bgra
mp = n * n / 1e6
w, h, gdim, rows = P.band_layout(data)
if rank == 0:
    print("encode: %.0f MP in %.2f s = %.0f MP/s, %.1f MB, %.3f bpp; %d group rows of %d px" % (mp, t_enc, mp / t_enc, len(data) / 1e6, len(data) * 8 / n / n, rows, gdim), flush=True)
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
