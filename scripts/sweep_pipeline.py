"""Pipelined sub-batches: a step of B images is submitted as `parts` asynchronous sub-batches (JxlB200DecodeBatchSubmit) with at most
`depth` of them in flight; compares with the synchronous call. Device-resident inputs and outputs, one GPU.
Usage: python scripts/sweep_pipeline.py [batch] [parts:depth:in_flight_per_part ...]"""
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
def submit(idx, infl):
    return P.decode_batch_submit(None, device=0, max_in_flight=infl, device_inputs=[dev_in[i].data_ptr() for i in idx], device_outputs=[dev_out[i].data_ptr() for i in idx],
                                 sizes=[dev_in[i].numel() for i in idx], out_sizes=[W * H * 3] * len(idx))
def run(steps, parts, depth, infl):
    chunks = [list(range(B * k // parts, B * (k + 1) // parts)) for k in range(parts)]
    pending = []
    for _ in range(steps):
        for c in chunks:
            pending.append(submit(c, infl))
            while len(pending) > depth - 1 and depth > 1 and len(pending) >= depth:
                assert all(s == 0 for s in pending.pop(0).wait())
            if depth == 1:
                assert all(s == 0 for s in pending.pop(0).wait())
    while pending:
        assert all(s == 0 for s in pending.pop(0).wait())
configs = [(1, 1, 128), (2, 2, 64), (4, 2, 64), (4, 3, 48), (2, 2, 96), (8, 3, 48)] if len(sys.argv) < 3 else [tuple(int(v) for v in c.split(":")) for c in sys.argv[2:]]
for parts, depth, infl in configs:
    run(2, parts, depth, infl)
    torch.cuda.synchronize(); steps = 4; t = time.time(); run(steps, parts, depth, infl); torch.cuda.synchronize(); dt = (time.time() - t) / steps
    print(json.dumps({"parts": parts, "depth": depth, "in_flight_per_part": infl, "batch": B, "mp_s": round(B * W * H / 1e6 / dt), "ms_per_step": round(dt * 1e3, 1)}), flush=True)
