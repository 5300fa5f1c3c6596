"""Sweeps the AC sections-per-warp setting and the in-flight depth of JxlB200DecodeBatch (device-resident in/out).
Usage: python scripts/sweep_lanes.py [batch] ; prints MP/s per (lanes, in_flight)."""
import os, sys, time, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, torch, pkgload, synth
    P = pkgload.load()
    B, inflight = int(sys.argv[2]), int(sys.argv[3]); W, H = 4000, 3000
    files = []
    for s in range(4):
        img = synth.synthetic_image(W, H, seed=s); bgra = np.concatenate([img[..., ::-1], np.full((H, W, 1), 255, np.uint8)], axis=2)
        files.append(P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=7)))
    files = [files[i % 4] for i in range(B)]
    dev_in = [torch.frombuffer(bytearray(f), dtype=torch.uint8).cuda() for f in files]; dev_out = [torch.empty(W * H * 3, dtype=torch.uint8, device="cuda") for _ in range(B)]
    def step():
        st = P.decode_batch(None, device=0, max_in_flight=inflight, device_inputs=[t.data_ptr() for t in dev_in], device_outputs=[t.data_ptr() for t in dev_out], sizes=[t.numel() for t in dev_in], out_sizes=[W * H * 3] * B)
        assert all(s == 0 for s in st)
    for _ in range(3): step()
    torch.cuda.synchronize(); n = 6; ts = []
    for _ in range(n):
        t = time.time(); step(); torch.cuda.synchronize(); ts.append(time.time() - t)
    dt = sum(ts) / n
    print("step ms:", [round(x * 1e3) for x in ts])
    print(json.dumps({"lanes": os.environ.get("JXLB200_AC_LANES"), "in_flight": inflight, "batch": B, "mp_s": B * W * H / 1e6 / dt, "ms_per_image": dt * 1e3 / B}))
    sys.exit(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for lanes, infl in [(8, 128), (8, 192), (8, 256), (16, 256), (4, 256)]:
    env = dict(os.environ, JXLB200_AC_LANES=str(lanes))
    r = subprocess.run([sys.executable, __file__, "child", str(B), str(infl)], env=env, capture_output=True, text=True)
    print("\n".join(r.stdout.strip().splitlines()[-2:]) if r.stdout.strip() else "FAILED " + r.stderr[-400:], flush=True)
