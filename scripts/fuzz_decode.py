"""Robustness sweep of LoadImage on mutated files: bit flips, byte overwrites and truncations anywhere in the file (headers AND section data,
i.e. the on-device entropy decoders see garbage). Every call must come back with a status — a crash or a hang fails the sweep (run it under
`timeout`). Prints the status histogram. Usage: python scripts/fuzz_decode.py [seconds] [--host-only]"""
import os, random, sys, time, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pkgload, oracle_py as O, spec_cases
P = pkgload.load()
budget = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 30.0
host_only = "--host-only" in sys.argv
img = O.synthetic_image(200, 136, seed=1, channels=4)
seeds = [O.encode(img, effort=7), O.encode(img, lossless=1), O.encode(img[..., :3], effort=3, use_prefix=1), O.encode(img, effort=5, num_passes=2),
         O.encode_layers(200, 136, [(img, {}), (img[:60, :80], dict(x0=15, y0=25, mode="blend"))], lossless=1)]
seeds += [bytes(c[1]) for c in spec_cases.cases() if c[0] in ("rgb8_palette_groups", "rgb8_lz77_multigroup", "rgb8_prev_channel_props", "rgb8_permuted_toc", "rgb8_palette_deltas",
                                                            "rgb8_local_trees_in_groups", "rgb8_palette_all_local", "rgba8_rct_per_group_local_trees", "rgba8_group_alpha_palettes",
                                                            "rgb8_group_rgb_palettes", "layers_three_slots_mul", "rgba8_two_palettes", "rgb8_squeeze_default_groups",
                                                            "rgb8_squeeze_lf_sections_local_trees", "rgba8_squeeze_explicit", "rgb8_squeeze_prev_channel_tree")]
rng = random.Random(4321); n = 0; t0 = time.time(); hist = {}
while time.time() - t0 < budget:
    s = bytearray(rng.choice(seeds))
    for _ in range(rng.randint(1, 4)):
        if len(s) < 10:
            break
        m = rng.random()
        if m < 0.55:
            s[rng.randrange(len(s))] ^= 1 << rng.randrange(8)
        elif m < 0.8:
            s[rng.randrange(len(s))] = rng.randrange(256)
        elif m < 0.9:
            del s[rng.randrange(8, len(s)):]
        else:
            a = rng.randrange(len(s)); s[a:a + 4] = bytes(rng.randrange(256) for _ in range(4))
    data = bytes(s)
    try:
        if host_only:
            P.peek_info(data); P.band_layout(data)
        else:
            image = P.DecoderImage(); P.JpegXLNative.LoadImage(data, image)
        hist["Ok"] = hist.get("Ok", 0) + 1
    except P.FormatException as e:
        hist[e.status] = hist.get(e.status, 0) + 1
    n += 1
print("%d mutants in %.0f s, every call returned; statuses: %s" % (n, time.time() - t0, dict(sorted(hist.items()))), flush=True)
