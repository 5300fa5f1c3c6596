"""CPU test: a short, seeded mutation sweep over the host-only entry points (JxlB200PeekInfo, JxlB200BandLayout, JxlB200DebugSectionSizes).
Every call must return a status; a crash or a hang of the header / TOC / box parsers fails the test. The GPU counterpart over whole decodes is
scripts/fuzz_decode.py (profiles/r02_fuzz.log)."""
import random
import time

import spec_cases


def test_mutated_files_never_crash_the_host_parsers(pkg, oracle):
    img = oracle.synthetic_image(64, 48, seed=1, channels=4)
    seeds = [oracle.encode(img, effort=7), oracle.encode(img, lossless=1), oracle.encode(img[..., :3], effort=3, exif=b"\0\0\0\0II*\0", xmp=b"<x/>"),
             oracle.encode_layers(64, 48, [(img, {}), (img[:20, :20], dict(x0=5, y0=5, mode="blend"))], lossless=1)]
    seeds += [bytes(c[1]) for c in spec_cases.cases()[:10]]
    seeds += [bytes(c[1]) for c in spec_cases.containerised()[:3]]
    rng = random.Random(20251018)
    statuses, n, t0 = {}, 0, time.time()
    while n < 20000 and time.time() - t0 < 5.0:
        s = bytearray(rng.choice(seeds))
        for _ in range(rng.randint(1, 4)):
            if len(s) < 10:
                break
            m = rng.random()
            if m < 0.6:
                s[rng.randrange(min(len(s), 256))] ^= 1 << rng.randrange(8)
            elif m < 0.8:
                s[rng.randrange(len(s))] = rng.randrange(256)
            elif m < 0.9:
                del s[rng.randrange(8, len(s)):]
            else:
                a = rng.randrange(len(s))
                s[a:a + 8] = bytes([0xff] * 8)          # e.g. a 64-bit box size of all ones
        data = bytes(s)
        for fn in (pkg.peek_info, pkg.band_layout, pkg.section_sizes):
            try:
                fn(data)
                statuses["Ok"] = statuses.get("Ok", 0) + 1
            except pkg.FormatException as e:
                statuses[e.status] = statuses.get(e.status, 0) + 1
        n += 1
    assert n >= 1000 and statuses.get("DecodeError", 0) > 0 and statuses.get("Ok", 0) > 0, (n, statuses)
