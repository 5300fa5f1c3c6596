"""Regenerates the golden fixtures in this directory with the CPU oracle (tests/oracle_py.py).

PARITY UNPINNED: the reference ships no .jxl files, golden vectors or tests and libjxl is unavailable offline,
so these files pin the oracle encoder/decoder pair against regressions and give the CUDA decoder fixed inputs; they are not
libjxl outputs. Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as O  # noqa: E402

CASES = [
    ("lossy_e3_64x48", dict(w=64, h=48, ch=3, seed=1), dict(effort=3)),
    ("lossy_e7_264x200", dict(w=264, h=200, ch=3, seed=2), dict(effort=7)),                       # gab + EPF + varblocks + CfL, 2 groups
    ("lossy_d3_rgba_130x90", dict(w=130, h=90, ch=4, seed=3), dict(effort=7, distance=3.0)),      # alpha, EPF x2, ragged size
    ("lossy_gray_72x56", dict(w=72, h=56, ch=1, seed=4), dict(effort=5)),
    ("lossy_prefix_dct16_96x96", dict(w=96, h=96, ch=3, seed=5), dict(effort=3, use_prefix=1, force_strategy=4)),
    ("lossless_rgba_70x50", dict(w=70, h=50, ch=4, seed=6), dict(lossless=1)),
    ("lossless_gray_40x30", dict(w=40, h=30, ch=1, seed=7), dict(lossless=1)),
    ("lossless_rgb_300x270", dict(w=300, h=270, ch=3, seed=8), dict(lossless=1, modular_group_shift=0)),   # 128x128 Modular groups
]


def main():
    index = {}
    for name, img_kw, enc_kw in CASES:
        img = O.synthetic_image(img_kw["w"], img_kw["h"], seed=img_kw["seed"], channels=img_kw["ch"])
        data = O.encode(img, **enc_kw)
        dec = O.decode(data)
        open(os.path.join(HERE, name + ".jxl"), "wb").write(data)
        np.save(os.path.join(HERE, name + ".npy"), dec.pixels)
        index[name] = dict(image=img_kw, encode=enc_kw, bytes=len(data), lossless=bool(enc_kw.get("lossless", 0)),
                           psnr=round(O.psnr(dec.pixels[..., :min(3, img.shape[2])], img[..., :min(3, img.shape[2])]), 3))
    json.dump(index, open(os.path.join(HERE, "index.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(index, indent=1))


if __name__ == "__main__":
    main()
