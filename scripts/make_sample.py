"""Encode a synthetic image with the engine's own GPU encoder (SaveImage path) and write it to disk."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, pkgload
from synth import synthetic_image
P = pkgload.load()
w, h, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
img = synthetic_image(w, h, seed=0)
bgra = np.concatenate([img[..., ::-1], np.full((h, w, 1), 255, np.uint8)], axis=2)
data = P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=7))
open(out, 'wb').write(data); print(len(data), 'bytes', len(data) * 8 / w / h, 'bpp')
