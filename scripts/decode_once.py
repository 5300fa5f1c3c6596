"""Decode one .jxl file a few times through the C ABI (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pkgload
P = pkgload.load()
data = open(sys.argv[1], 'rb').read()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for _ in range(n):
    out = P.load_image_bgra(data)
print(out.shape, P.last_stage_times())
