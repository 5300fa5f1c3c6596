"""Step time of JxlB200DecodeBatch against the batch size (device-resident in/out, one GPU): the slope is the steady-state cost per
image, the intercept the fill/drain latency of the three-phase pipeline. Usage: python scripts/sweep_batch_size.py [sizes...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, pkgload, synth
P = pkgload.load()
sizes = [int(v) for v in sys.argv[1:]] or [32, 64, 128, 256, 512]
W, H = 4000, 3000
base = []
for s in range(4):
    img = synth.synthetic_image(W, H, seed=s); bgra = np.concatenate([img[..., ::-1], np.full((H, W, 1), 255, np.uint8)], axis=2)
    base.append(P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=7)))
BM = max(sizes)
dev_in = [torch.frombuffer(bytearray(base[i % 4]), dtype=torch.uint8).cuda() for i in range(BM)]; dev_out = [torch.empty(W * H * 3, dtype=torch.uint8, device="cuda") for _ in range(BM)]
def step(B):
    st = P.decode_batch(None, device=0, max_in_flight=128, device_inputs=[t.data_ptr() for t in dev_in[:B]], device_outputs=[t.data_ptr() for t in dev_out[:B]], sizes=[t.numel() for t in dev_in[:B]], out_sizes=[W * H * 3] * B)
    assert all(s == 0 for s in st)
for _ in range(4): step(256)
for B in sizes:
    for _ in range(2): step(B)
    torch.cuda.synchronize(); ts = []
    for _ in range(4):
        t = time.time(); step(B); torch.cuda.synchronize(); ts.append(time.time() - t)
    dt = min(ts)
    print(json.dumps({"batch": B, "step_ms": round(dt * 1e3, 1), "ms_per_image": round(dt * 1e3 / B, 3), "mp_s": round(B * W * H / 1e6 / dt)}), flush=True)
