"""The single-image configurations of BASELINE.json (1, 2, 4a, 4b) through the reference-facing calls on one GPU, with the CPU oracle timed beside
them on the box's host cores (own restatement, NOT libjxl). Wall time of the call (host buffers, H2D + D2H inside) and the engine's CUDA-event
stage times. Configs 3 (batch) and 5 (gigapixel) are bench.py and scripts/gigapixel_bands.py. Usage: python scripts/time_configs.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pkgload, oracle_py as O
P = pkgload.load()
cores = os.cpu_count() or 1


def run(name, data, call, mp):
    call(data)                                                     # warm-up (first-use tables, pools)
    best, stages = 1e9, None
    for _ in range(3):
        t = time.time(); call(data); dt = (time.time() - t) * 1e3
        if dt < best:
            best, stages = dt, P.last_stage_times()
    t = time.time(); O.decode(data, threads=cores); cpu = (time.time() - t) * 1e3
    st = " ".join("%s %.2f" % (k, stages[k]) for k in ("lf", "ac", "recon", "filters", "output", "d2h", "total"))
    print("%-58s %8.2f ms wall = %7.0f MP/s | device ms: %s | CPU oracle, %d threads: %.0f ms = %.1f MP/s" % (name, best, mp / best * 1e3, st, cores, cpu, mp / cpu * 1e3), flush=True)


def load(data):
    image = P.DecoderImage(); P.JpegXLNative.LoadImage(data, image); return image


img = O.synthetic_image(1024, 768, seed=1)
run("1  1024x768 RGB8 VarDCT d=1 e=7, LoadImage", O.encode(img, effort=7, distance=1.0), load, 1024 * 768 / 1e6)
img = O.synthetic_image(3840, 2160, seed=2)
run("2  3840x2160 RGB8 VarDCT d=1 + BGRA pack, LoadImageBgra", O.encode(img, effort=7, distance=1.0, threads=cores), P.load_image_bgra, 3840 * 2160 / 1e6)
img = O.synthetic_image(4096, 4096, seed=3, channels=4)
run("4a 4096x4096 RGBA8 lossless Modular, LoadImage", O.encode(img, lossless=1, threads=cores), load, 4096 * 4096 / 1e6)
img = O.synthetic_image(7680, 4320, seed=0).astype(np.float32) / 255.0
for kw, nm in ((dict(bits=16), "16-bit"), (dict(bits=32, exp_bits=8), "float32")):
    data = O.encode(img, effort=3, gab=1, epf=2, primaries=9, tf=16, intensity_target=1000.0, threads=cores, **kw)
    run("4b 7680x4320 Rec.2020 PQ %s VarDCT, gaborish + EPFx2, LoadImage" % nm, data, load, 7680 * 4320 / 1e6)
