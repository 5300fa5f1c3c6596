"""Sweeps the number of host enqueue threads (JXLB200_HOST_THREADS) and the in-flight depth of JxlB200DecodeBatch on one GPU,
device-resident inputs and outputs (bench.py's `value` configuration). Usage: python scripts/sweep_host_threads.py [batch]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, pkgload, synth
P = pkgload.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H = 4000, 3000
files = []
for s in range(4):
    img = synth.synthetic_image(W, H, seed=s); bgra = np.concatenate([img[..., ::-1], np.full((H, W, 1), 255, np.uint8)], axis=2)
    files.append(P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=7)))
files = [files[i % 4] for i in range(B)]
dev_in = [torch.frombuffer(bytearray(f), dtype=torch.uint8).cuda() for f in files]; dev_out = [torch.empty(W * H * 3, dtype=torch.uint8, device="cuda") for _ in range(B)]
def step(infl):
    st = P.decode_batch(None, device=0, max_in_flight=infl, device_inputs=[t.data_ptr() for t in dev_in], device_outputs=[t.data_ptr() for t in dev_out], sizes=[t.numel() for t in dev_in], out_sizes=[W * H * 3] * B)
    assert all(s == 0 for s in st)
configs = [(1, 128), (2, 128), (4, 128), (8, 128), (4, 256), (8, 256)] if len(sys.argv) < 3 else [tuple(int(v) for v in c.split(":")) for c in sys.argv[2:]]
for cfg in configs:
    threads, infl = cfg[0], cfg[1]
    os.environ["JXLB200_HOST_THREADS"] = str(threads)
    os.environ["JXLB200_AC_LANES"] = str(cfg[2]) if len(cfg) > 2 else "0"
    os.environ["JXLB200_BUNDLE"] = str(cfg[3]) if len(cfg) > 3 else "0"
    os.environ["JXLB200_AC_SMEM_KB"] = str(cfg[4]) if len(cfg) > 4 else "0"
    for _ in range(3): step(infl)
    torch.cuda.synchronize(); ts = []
    for _ in range(4):
        t = time.time(); step(infl); torch.cuda.synchronize(); ts.append(time.time() - t)
    dt = sum(ts) / len(ts)
    print(json.dumps({"host_threads": threads, "in_flight": infl, "lanes": os.environ["JXLB200_AC_LANES"], "bundle": os.environ["JXLB200_BUNDLE"], "ac_smem_kb": os.environ["JXLB200_AC_SMEM_KB"], "batch": B, "mp_s": round(B * W * H / 1e6 / dt), "ms_per_image": round(dt * 1e3 / B, 3), "step_ms": [round(x * 1e3) for x in ts]}), flush=True)
