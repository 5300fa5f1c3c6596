"""End-to-end batch decode with HOST buffers (files in pageable host memory, pixels into page-locked host memory): synchronous call against
asynchronous sub-batches with `depth` in flight, where the D2H of one sub-batch overlaps the entropy phases of the next.
Usage: python scripts/sweep_e2e.py [batch] [parts:depth:in_flight ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, pkgload, synth
P = pkgload.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H = 4000, 3000
base = []
for s in range(4):
    img = synth.synthetic_image(W, H, seed=s); bgra = np.concatenate([img[..., ::-1], np.full((H, W, 1), 255, np.uint8)], axis=2)
    base.append(P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=7)))
files = [base[i % 4] for i in range(B)]
host = [torch.empty(W * H * 3, dtype=torch.uint8).pin_memory() for _ in range(B)]; host_np = [t.numpy() for t in host]
def run(steps, parts, depth, infl):
    chunks = [list(range(B * k // parts, B * (k + 1) // parts)) for k in range(parts)]
    pending = []
    for _ in range(steps):
        for c in chunks:
            pending.append(P.decode_batch_submit([files[i] for i in c], [host_np[i] for i in c], device=0, max_in_flight=infl))
            while len(pending) >= depth:
                assert all(s == 0 for s in pending.pop(0).wait())
    while pending:
        assert all(s == 0 for s in pending.pop(0).wait())
configs = [(1, 1, 128), (2, 2, 128), (2, 2, 64), (3, 2, 64), (4, 2, 64), (4, 3, 64)] if len(sys.argv) < 3 else [tuple(int(v) for v in c.split(":")) for c in sys.argv[2:]]
run(3, 1, 1, 128)
for parts, depth, infl in configs:
    run(2, parts, depth, infl)
    torch.cuda.synchronize(); steps = 4; t = time.time(); run(steps, parts, depth, infl); torch.cuda.synchronize(); dt = (time.time() - t) / steps
    print(json.dumps({"parts": parts, "depth": depth, "in_flight_per_part": infl, "batch": B, "e2e_mp_s": round(B * W * H / 1e6 / dt), "ms_per_step": round(dt * 1e3, 1)}), flush=True)
