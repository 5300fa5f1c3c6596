"""Encode a synthetic image with the CPU oracle's encoder using the bench's "mixed" block pattern and write it to disk (profiling input)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O
w, h, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
img = O.synthetic_image(w, h, seed=0)
data = O.encode(img, effort=7, distance=1.0, varblock_scale=4.0, varblock_pattern=1, threads=os.cpu_count() or 1)
open(out, "wb").write(data); print(len(data), "bytes", O.last_encode_strategy_cells())
