"""SaveImage at efforts 3 / 5 / 7 on the same surface: file size and PSNR of the oracle's decode of each file (what adaptive quantisation and
chroma from luma, on from effort 5, buy). Usage: python scripts/compare_efforts.py [w h]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pkgload, oracle_py as O
P = pkgload.load()
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 768)
for seed in (1, 2):
    img = O.synthetic_image(w, h, seed=seed)
    bgra = np.ascontiguousarray(np.concatenate([img[..., ::-1], np.full((h, w, 1), 255, np.uint8)], axis=2))
    for q in (90, 75):
        for effort, aq, cfl in ((3, 1, 1), (5, 0, 0), (5, 0, 1), (5, 1, 0), (5, 1, 1)):
            os.environ["JXLB200_ENC_AQ"], os.environ["JXLB200_ENC_CFL"] = str(aq), str(cfl)
            data = P.encode_to_memory(bgra, P.EncoderOptions(quality=q, effort=effort))
            back = O.decode(data, threads=8).pixels
            tools = "" if effort < 5 else " (gaborish + EPF%s%s)" % (", adaptive quantisation" if aq else "", ", chroma from luma" if cfl else "")
            print("seed %d quality %d effort %d%s: %7d bytes  %.3f bpp  PSNR %.2f dB" % (seed, q, effort, tools, len(data), len(data) * 8 / (w * h), O.psnr(back, img)), flush=True)
