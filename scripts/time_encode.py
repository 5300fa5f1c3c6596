"""SaveImage-path timing on one GPU: a 4000x3000 BGRA surface encoded with JxlB200EncodeToMemory (host surface, H2D included) at the
efforts / modes given; prints wall ms, device ms (the engine's own CUDA-event total) and MP/s. Usage: python scripts/time_encode.py [--reps 3]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
from synth import synthetic_image

ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--w", type=int, default=4000); ap.add_argument("--h", type=int, default=3000)
args = ap.parse_args()
P = pkgload.load()
img = synthetic_image(args.w, args.h, seed=1, channels=4)
bgra = np.ascontiguousarray(np.concatenate([img[..., 2::-1], img[..., 3:]], axis=2))
opaque = bgra.copy(); opaque[..., 3] = 255
mp = args.w * args.h / 1e6
for name, surf, opts in (("lossy d=1.0 e=1 RGB (prefix codes)", opaque, P.EncoderOptions(quality=90, effort=1)), ("lossy d=1.0 e=3 RGB", opaque, P.EncoderOptions(quality=90, effort=3)), ("lossy d=1.0 e=7 RGB (gaborish)", opaque, P.EncoderOptions(quality=90, effort=7)),
                         ("lossy d=1.0 e=3 RGBA", bgra, P.EncoderOptions(quality=90, effort=3)), ("lossless RGBA", bgra, P.EncoderOptions(lossless=True)), ("lossless RGBA e=1 (prefix codes)", bgra, P.EncoderOptions(lossless=True, effort=1))):
    P.encode_to_memory(surf, opts)                         # warm-up: tables, first-use allocations
    best, dev, size = 1e9, 0.0, 0
    for _ in range(args.reps):
        t = time.time(); data = P.encode_to_memory(surf, opts); dt = (time.time() - t) * 1e3
        if dt < best:
            best, dev, size = dt, P.last_stage_times()["total"], len(data)
    print("%-34s wall %7.1f ms = %6.0f MP/s   device sections %7.1f ms   %.3f bpp" % (name, best, mp / best * 1e3, dev, size * 8 / (args.w * args.h)), flush=True)
