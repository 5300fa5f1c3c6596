"""Small ICC helpers for the tests: build a matrix/TRC profile, parse one back (independent of the engine's C++ code)."""
import struct
import numpy as np

D50 = (0.9642, 1.0, 0.8249)
BRADFORD = np.array([[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]])


def s15(v):
    return struct.pack(">i", int(round(v * 65536.0)))


def adapt_to_d50(wx, wy):
    w = np.array([wx / wy, 1.0, (1 - wx - wy) / wy])
    lw, ld = BRADFORD @ w, BRADFORD @ np.array(D50)
    return np.linalg.inv(BRADFORD) @ np.diag(ld / lw) @ BRADFORD


def rgb_to_xyz(prims, wx, wy):
    P = np.array([[p[0] / p[1] for p in prims], [1.0] * 3, [(1 - p[0] - p[1]) / p[1] for p in prims]])
    W = np.array([wx / wy, 1.0, (1 - wx - wy) / wy])
    return P * (np.linalg.inv(P) @ W)[None, :]


SRGB_PRIMS = ((0.639998686, 0.330010138), (0.300003784, 0.600003357), (0.150002046, 0.059997204))
P3_PRIMS = ((0.680, 0.320), (0.265, 0.690), (0.150, 0.060))
ADOBE_PRIMS = ((0.64, 0.33), (0.21, 0.71), (0.15, 0.06))


def make_matrix_icc(prims, white=(0.3127, 0.3290), gamma=2.2, curve="para"):
    """ICC v4 matrix/TRC display profile: desc, wtpt, chad, r/g/bXYZ, shared TRC ('para' type 0, or 'curv' with one gamma entry / a table)."""
    chad = adapt_to_d50(*white)
    m = chad @ rgb_to_xyz(prims, *white)
    text = "test profile"
    desc = b"mluc" + b"\0" * 4 + struct.pack(">III", 1, 12, 0x656E5553) + struct.pack(">II", len(text) * 2, 28) + text.encode("utf-16-be")
    tags = [(b"desc", desc), (b"wtpt", b"XYZ " + b"\0" * 4 + b"".join(s15(v) for v in D50)),
            (b"chad", b"sf32" + b"\0" * 4 + b"".join(s15(v) for v in chad.reshape(-1)))]
    for k, sig in enumerate((b"rXYZ", b"gXYZ", b"bXYZ")):
        tags.append((sig, b"XYZ " + b"\0" * 4 + b"".join(s15(m[r, k]) for r in range(3))))
    if curve == "para":
        trc = b"para" + b"\0" * 4 + struct.pack(">HH", 0, 0) + s15(gamma)
    elif curve == "curv1":
        trc = b"curv" + b"\0" * 4 + struct.pack(">I", 1) + struct.pack(">H", int(round(gamma * 256)))
    else:   # sampled table
        n = 1024
        trc = b"curv" + b"\0" * 4 + struct.pack(">I", n) + b"".join(struct.pack(">H", int(round(65535 * (i / (n - 1)) ** gamma))) for i in range(n))
    for sig in (b"rTRC", b"gTRC", b"bTRC"):
        tags.append((sig, trc))
    base = 128 + 4 + 12 * len(tags)
    body, table, prev = b"", b"", None
    for sig, data in tags:
        if prev is not None and prev[1] == data and sig[1:] == b"TRC":
            off, ln = prev[2], prev[3]
        else:
            body += b"\0" * (-len(body) % 4)
            off, ln = base + len(body), len(data)
            body += data
        table += sig + struct.pack(">II", off, ln)
        prev = (sig, data, off, ln)
    body += b"\0" * (-len(body) % 4)
    size = base + len(body)
    hdr = struct.pack(">I", size) + b"test" + struct.pack(">I", 0x04400000) + b"mntr" + b"RGB " + b"XYZ " + struct.pack(">6H", 2024, 1, 1, 0, 0, 0)
    hdr += b"acsp" + b"APPL" + b"\0" * 20 + struct.pack(">I", 0) + b"".join(s15(v) for v in D50) + b"test"
    hdr += b"\0" * (128 - len(hdr))
    return hdr + struct.pack(">I", len(tags)) + table + body


def parse_icc(icc):
    """Returns dict: size, version, cls, space, pcs, intent, tags {sig: bytes}."""
    size = struct.unpack(">I", icc[0:4])[0]
    n = struct.unpack(">I", icc[128:132])[0]
    tags = {}
    for i in range(n):
        sig = icc[132 + 12 * i: 136 + 12 * i]
        off, ln = struct.unpack(">II", icc[136 + 12 * i: 144 + 12 * i])
        assert off + ln <= len(icc)
        tags[sig] = icc[off: off + ln]
    return {"size": size, "version": icc[8], "cls": icc[12:16], "space": icc[16:20], "pcs": icc[20:24], "intent": struct.unpack(">I", icc[64:68])[0], "tags": tags}


def xyz_of(tag):
    assert tag[:4] == b"XYZ "
    return np.array([struct.unpack(">i", tag[8 + 4 * i: 12 + 4 * i])[0] / 65536.0 for i in range(3)])


def para_of(tag):
    assert tag[:4] == b"para"
    t = struct.unpack(">H", tag[8:10])[0]
    n = {0: 1, 1: 3, 2: 4, 3: 5, 4: 7}[t]
    return t, [struct.unpack(">i", tag[12 + 4 * i: 16 + 4 * i])[0] / 65536.0 for i in range(n)]
