"""Host ceiling of the end-to-end batch metric: every rank copies its share of one step's decoded pixels (36 MB per 12 MP RGB8 image)
from device to page-locked host memory, all ranks at once, nothing else running. bench.py's e2e figure cannot exceed
(pixels per step) / (this time); the driver's 8-GPU e2e efficiency is bounded by the host, not by the decode path, when the two agree.
Usage: [torchrun --nproc-per-node N] python scripts/d2h_ceiling.py [--batch 256] [--scaling strong|weak]"""
import argparse, json, os, sys, time
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=256); ap.add_argument("--scaling", default="strong"); ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 4000, 3000
n = args.batch // world if args.scaling == "strong" else args.batch
one = W * H * 3
dev = [torch.empty(one, dtype=torch.uint8, device="cuda") for _ in range(n)]
host = [torch.empty(one, dtype=torch.uint8).pin_memory() for _ in range(n)]
streams = [torch.cuda.Stream() for _ in range(4)]
def step():
    for i in range(n):
        with torch.cuda.stream(streams[i % 4]):
            host[i].copy_(dev[i], non_blocking=True)
    torch.cuda.synchronize()
for _ in range(2): step()
if dist is not None: dist.barrier()
t = time.time()
for _ in range(args.steps): step()
dt = (time.time() - t) / args.steps
if dist is not None:
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
if rank == 0:
    total = n * world * one
    print(json.dumps({"what": "concurrent D2H of one step's pixels into page-locked host memory", "n_gpus": world, "scaling": args.scaling, "images_per_step": n * world,
                      "bytes_per_step": total, "ms_per_step": dt * 1e3, "aggregate_gb_s": total / dt / 1e9, "e2e_ceiling_mp_s": n * world * W * H / 1e6 / dt, "host_cpus": os.cpu_count()}))
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
