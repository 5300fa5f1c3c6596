#!/usr/bin/env python
"""bench.py — decode MP/s of lossy VarDCT 12 MP 8-bit images (BASELINE.json metric) on N B200s of one node.

A step = one pass of the hot path (LoadImage-equivalent decode) over one batch of synthetic 4000x3000 RGB8
VarDCT d=1.0 files per GPU (BASELINE config 3, files sharded across ranks, no data-path collective: weak scaling).
  value : whole-job MP/s with the compressed files already resident in HBM and the decoded pixels left in HBM
  e2e   : the same through the C ABI with HOST buffers (pinned), H2D of the files and D2H of the pixels inside the timed region
  roofline      : the kernel with the largest share of a decode, algorithmic bytes / CUDA-event duration vs measured HBM peak
  stage_rooflines: the HBM-bound stage kernels (dequant+IDCT, gaborish+EPF, colour+pack) — the stages north_star sets the 50 % bar on
  cpu_baseline  : the CPU oracle (scalar port, NOT libjxl — libjxl is unavailable offline) on a bounded sample, all host cores
`--impl reference` times that CPU path alone (the reference's libjxl path cannot be built here: N/vcpkg.json:6-9 is un-vendored).
"""
import argparse
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per in-flight image stream (default 8 serialises streams)
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decode MP/s (lossy VarDCT, 12 MP 8-bit)"
UNIT = "MP/s"


WORKLOAD = ("batch of %d synthetic %dx%d RGB8 VarDCT d=1.0 e=7 files per GPU (%d distinct seeds, encoded by the engine's SaveImage path), "
            "decoded to interleaved RGB8; files sharded across ranks, no collective")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--width", type=int, default=4000)
    ap.add_argument("--height", type=int, default=3000)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic images (seeds 0..distinct-1), repeated to fill the batch")
    ap.add_argument("--in-flight", type=int, default=128)
    ap.add_argument("--cpu-sample", type=int, default=2, help="images in the bounded CPU-baseline sample")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        # In-process NVML when available (a clock read costs microseconds); spawning nvidia-smi five times a second would steal
        # host time from the single enqueue thread being measured. Falls back to nvidia-smi.
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop_evt.is_set():
                r = int(get_reasons(h))
                act = lambda bit: "Active" if r & bit else "Not Active"
                self.rows.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)])
                self._stop_evt.wait(0.1)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def make_files(P, args, rank):
    """Synthetic images -> .jxl files with the engine's own SaveImage path (quality 90 -> d=1.0, effort 7). Setup, not timed."""
    import numpy as np
    from synth import synthetic_image
    files = []
    for seed in range(args.distinct):
        img = synthetic_image(args.width, args.height, seed=seed + 1000 * rank)
        bgra = np.concatenate([img[..., ::-1], np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2)
        files.append(P.encode_to_memory(bgra, P.EncoderOptions(quality=90, effort=7)))
    return [files[i % len(files)] for i in range(args.batch)]


def cpu_baseline(files, args, sample):
    """CPU oracle (tests/oracle_py -> oracle/_build/liboracle.so) on a bounded sample with every host core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    cores = os.cpu_count() or 1
    O.decode(files[0], threads=cores)   # warm-up (page in the library, cosine tables)
    t0 = time.perf_counter()
    for i in range(sample):
        O.decode(files[i % len(files)], threads=cores)
    dt = time.perf_counter() - t0
    mp = sample * args.width * args.height / 1e6
    return {"value": mp / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d images of %dx%d decoded by the scalar CPU oracle (own restatement, NOT libjxl) with %d threads over groups" % (sample, args.width, args.height, cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pkgload
    files = None
    try:
        P = pkgload.load(build_if_missing=True)
        ok, _ = P.cuda_available()
        if ok:
            files = make_files(P, args, 0)
    except Exception:
        files = None
    if files is None:   # no GPU to run SaveImage on: use the oracle's own encoder for the inputs
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as O
        from synth import synthetic_image
        files = [O.encode(synthetic_image(args.width, args.height, seed=s), effort=7, distance=1.0, threads=os.cpu_count() or 1) for s in range(min(args.distinct, 2))]
    base = cpu_baseline(files, args, 1)   # warm
    per_step = args.cpu_sample
    for _ in range(args.warmup):
        cpu_baseline(files, args, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        base = cpu_baseline(files, args, per_step)
    dt = time.perf_counter() - t0
    mp_step = per_step * args.width * args.height / 1e6
    # cpu_baseline() includes one warm-up decode per call; report the timed sample rate it measured
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (args.batch, args.width, args.height, args.distinct), "mp_per_step": mp_step,
                       "sample": "the same files decoded on the host CPU; each step a bounded sample of %d images" % per_step},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(torch, local):
    """Pins this rank to the CPUs local to its GPU (what `numactl` would do), so that its page-locked buffers and its enqueue thread
    sit on the GPU's NUMA node. Silently does nothing where sysfs does not expose the topology."""
    if os.environ.get("JXLB200_BENCH_NOBIND"):
        return
    try:
        pr = torch.cuda.get_device_properties(local)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(torch, local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pkgload
    P = pkgload.load(build_if_missing=True)
    ok, why = P.cuda_available()
    if not ok:
        raise SystemExit("bench.py needs a GPU: " + why)
    files = make_files(P, args, rank)
    B, W, H = args.batch, args.width, args.height
    mp_step = B * W * H / 1e6
    comp_bytes = sum(len(f) for f in files)
    out_bytes_one = W * H * 3
    # device-resident inputs / outputs (value) and pinned host buffers (e2e)
    dev_in = [torch.frombuffer(bytearray(f), dtype=torch.uint8).cuda() for f in files]
    dev_out = [torch.empty(out_bytes_one, dtype=torch.uint8, device="cuda") for _ in range(B)]
    try:
        host_out = [torch.empty(out_bytes_one, dtype=torch.uint8).pin_memory() for _ in range(B)]
        host_pinned = True
    except RuntimeError:   # the box cannot page-lock B x 36 MB per rank: pageable buffers (the engine then stages through its own pinned pool)
        host_out = [torch.empty(out_bytes_one, dtype=torch.uint8) for _ in range(B)]
        host_pinned = False
    host_out_np = [t.numpy() for t in host_out]
    torch.cuda.synchronize()

    def step_device():
        st = P.decode_batch(None, device=local, max_in_flight=args.in_flight, device_inputs=[t.data_ptr() for t in dev_in], device_outputs=[t.data_ptr() for t in dev_out],
                            sizes=[t.numel() for t in dev_in], out_sizes=[out_bytes_one] * B)
        assert all(s == 0 for s in st)

    def step_host():
        st = P.decode_batch(files, host_out_np, device=local, max_in_flight=args.in_flight)
        assert all(s == 0 for s in st)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = P.kernel_launch_count()
    ms_dev = timed(step_device, args.steps)
    launches = P.kernel_launch_count() - launches0
    clocks = sampler.stop()
    for _ in range(max(1, args.warmup // 2)):
        step_host()
    ms_host = timed(step_host, args.steps)
    if dist is not None:
        lt = torch.tensor([launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())

    value = world * mp_step * args.steps / (ms_dev / 1e3)
    e2e = world * mp_step * args.steps / (ms_host / 1e3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (B, W, H, args.distinct),
                       "mp_per_step_per_gpu": mp_step, "bpp": 8.0 * comp_bytes / (B * W * H), "in_flight": args.in_flight,
                       "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % ((comp_bytes + B * out_bytes_one) / 1e6)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": comp_bytes, "d2h_bytes_per_step": B * out_bytes_one, "ms_per_step": ms_host / args.steps, "host_buffers": "page-locked" if host_pinned else "pageable"},
            "gpu_launches": launches, "clocks": clocks}

    if rank == 0:
        # per-kernel durations of one decode, CUDA events on the stream the kernels are launched on (engine StageTimes)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        acc = {}
        reps = 5
        for i in range(reps + 1):
            P.load_image_bgra(files[0]) if False else None
            img = P.DecoderImage()
            P.JpegXLNative.LoadImage(files[0], img)
            if i:   # first one is a warm-up
                for k, v in P.last_stage_times().items():
                    acc[k] = acc.get(k, 0.0) + v / reps
        px = W * H
        comp_px = len(files[0]) / px
        stage_bytes = {"lf": (comp_px * 0.12 + (12 + 3) / 64.0), "ac": (comp_px * 0.88 + 6.0), "recon": 18.0, "filters": 24.0 * 2, "output": 15.0}
        names = {"lf": "k_lf_group (LF coefficients + HF metadata entropy decode; latency-bound serial streams)", "ac": "k_ac_group (AC coefficient entropy decode; latency-bound serial streams)",
                 "recon": "k_reconstruct (dequant + CfL + IDCT)", "filters": "k_gaborish + k_epf<1>", "output": "k_output (XYB->sRGB + interleave)"}
        if acc.get("output", 0.0) < 0.02 and acc.get("filters", 0.0) > 0:   # fused path: one kernel reads XYB once and writes RGB8 once
            stage_bytes["filters"] = 15.0
            acc["filters"] += acc.get("output", 0.0)
            acc["output"] = 0.0
            names["filters"] = "k_render<GAB,EPF> (gaborish + EPF + XYB->sRGB + interleave fused; XYB read once, RGB8 written once)"
        stages = {}
        for k in ("lf", "ac", "recon", "filters", "output"):
            ms = acc.get(k, 0.0)
            if ms > 0:
                ach = stage_bytes[k] * px / (ms * 1e-3) / 1e9
                stages[k] = {"kernel": names[k], "ms": ms, "bytes_per_px": stage_bytes[k], "achieved": ach, "frac": ach / peak}
        dom = max(stages, key=lambda k: stages[k]["ms"]) if stages else None
        if dom:
            d = stages[dom]
            line["roofline"] = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved"], "peak": peak, "unit": "GB/s", "frac": d["frac"], "traffic": None,
                                "peak_source": peak_src, "share_of_decode": d["ms"] / max(acc.get("total", 1e-9), 1e-9),
                                "note": "dominant kernel of a single-image decode; entropy decode is a serial bit stream per section, so its HBM fraction is tiny by nature — ns/symbol is the relevant figure (DESIGN.md)"}
        # DRAM traffic per launch from the one `ncu --set full` capture of these kernels at this frame size (profiles/r01_ncu_render.md)
        if W * H == 4000 * 3000:
            if "recon" in stages:
                stages["recon"]["traffic"] = 74.7e6 + 90.1e6
            if "filters" in stages and stage_bytes["filters"] == 15.0:
                stages["filters"]["traffic"] = 144.9e6 + 36.1e6
        line["stage_rooflines"] = {k: v for k, v in stages.items() if k in ("recon", "filters", "output")}
        line["single_image_ms"] = acc
        try:
            line["cpu_baseline"] = cpu_baseline(files, args, args.cpu_sample)
        except Exception as e:   # the oracle library may be absent; the GPU numbers stand on their own
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "unavailable: %s" % e}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
