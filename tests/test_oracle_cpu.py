"""CPU tests of the oracle (the checker): known-answer material constructible offline and the committed goldens.
PARITY UNPINNED against libjxl — these pin the restatement against itself, closed forms and the reference's own constants."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
INDEX = json.load(open(os.path.join(GOLDEN, "index.json")))


@pytest.mark.parametrize("name", sorted(INDEX))
def test_golden_files_decode_to_committed_pixels(oracle, name):
    data = open(os.path.join(GOLDEN, name + ".jxl"), "rb").read()
    want = np.load(os.path.join(GOLDEN, name + ".npy"))
    got = oracle.decode(data, threads=2).pixels
    assert got.shape == want.shape and np.array_equal(got, want)
    meta = INDEX[name]
    if meta["lossless"]:
        src = oracle.synthetic_image(meta["image"]["w"], meta["image"]["h"], seed=meta["image"]["seed"], channels=meta["image"]["ch"])
        assert np.array_equal(got, src)


def test_golden_files_are_what_the_encoder_still_produces(oracle):
    for name in ("lossy_e3_64x48", "lossless_gray_40x30"):
        m = INDEX[name]
        img = oracle.synthetic_image(m["image"]["w"], m["image"]["h"], seed=m["image"]["seed"], channels=m["image"]["ch"])
        assert oracle.encode(img, **m["encode"]) == open(os.path.join(GOLDEN, name + ".jxl"), "rb").read()


@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 18, 19, 20, 21])
def test_transforms_invert_and_match_fp64_dct(oracle, strategy):
    import ctypes as C
    L = oracle.lib()
    cy = [1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16][strategy]
    cx = [1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16][strategy]
    H, W = cy * 8, cx * 8
    rng = np.random.default_rng(strategy)
    px = rng.standard_normal((H, W)).astype(np.float32)
    coef = np.zeros(H * W, np.float32)
    back = np.zeros((H, W), np.float32)
    L.jxlo_transform_from_pixels(strategy, px.ctypes.data, W, coef.ctypes.data)
    L.jxlo_transform_to_pixels(strategy, coef.ctypes.data, back.ctypes.data, W)
    assert np.max(np.abs(back - px)) < 2e-5
    if strategy in (0, 4, 5, 18, 21):   # square plain DCTs: compare with the textbook fp64 DCT-II scaled by 1/N per dimension
        def dct_mat(n):
            k, i = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
            m = np.cos((2 * i + 1) * k * np.pi / (2 * n)) * np.where(k == 0, 1.0, np.sqrt(2.0))
            return m / n
        F = dct_mat(H) @ px.astype(np.float64) @ dct_mat(W).T      # F[vf][hf]
        S = coef.reshape(W, H)                                      # square blocks are stored transposed: S[hf][vf]
        assert np.max(np.abs(S.T - F)) < 1e-5
        assert abs(S[0, 0] - px.mean()) < 1e-6                      # DC = block mean (A.9 scaling)


def test_resample_scale_closed_form(oracle):
    def s(n, u):
        return 1.0 / np.prod([math.cos(u * math.pi * (1 << k) / (16.0 * n)) for k in range(3)])
    assert abs(s(2, 1) - 1.108937353592731823) < 1e-12
    for u, v in zip((1, 2, 3), (1.02576009678, 1.10893735359, 1.27055936877)):
        assert abs(s(4, u) - v) < 1e-9
    # LLF of a DCT16 block from a constant LF patch is the constant itself in position 0
    import ctypes as C
    L = oracle.lib()
    dc = np.full((2, 2), 0.37, np.float32)
    block = np.zeros(256, np.float32)
    L.jxlo_llf_from_dc(4, dc.ctypes.data, 2, block.ctypes.data)
    assert abs(block[0] - 0.37) < 1e-6 and abs(block[1]) < 1e-6 and abs(block[16]) < 1e-6


def test_opsin_matrices_are_inverse_and_colour_roundtrips(oracle):
    M = np.array([[0.30, 0.622, 0.078], [0.23, 0.692, 0.078], [0.24342268924547819, 0.20476744424496821, 0.55180986650955360]])
    Mi = np.array([[11.031566901960783, -9.866943921568629, -0.16462299647058826], [-3.254147380392157, 4.418770392156863, -0.16462299647058826],
                   [-3.6588512862745097, 2.7129230470588235, 1.9459282392156863]])
    assert np.max(np.abs(Mi @ M - np.eye(3))) < 1e-6
    L = oracle.lib()
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (4096, 3), dtype=np.uint8)
    x, y, b = (np.zeros(4096, np.float32) for _ in range(3))
    L.jxlo_srgb8_to_xyb(rgb.ctypes.data, 4096, x.ctypes.data, y.ctypes.data, b.ctypes.data)
    out = np.zeros((4096, 3), np.uint8)
    L.jxlo_xyb_to_srgb8(x.ctypes.data, y.ctypes.data, b.ctypes.data, 4096, out.ctypes.data)
    assert np.array_equal(out, rgb)


def test_natural_order_is_a_permutation_with_llf_first(oracle):
    L = oracle.lib()
    sizes = [64, 64, 256, 1024, 128, 256, 512, 4096, 2048, 16384, 8192, 65536, 32768]
    for o, n in enumerate(sizes):
        buf = np.zeros(n, np.uint32)
        assert L.jxlo_natural_order(o, buf.ctypes.data) == n
        assert np.array_equal(np.sort(buf), np.arange(n, dtype=np.uint32))
    buf = np.zeros(64, np.uint32)
    L.jxlo_natural_order(0, buf.ctypes.data)
    assert list(buf[:4]) == [0, 1, 8, 16]


def test_dequant_tables_positive_and_dct8_values(oracle):
    L = oracle.lib()
    n = L.jxlo_dequant_table(0, None)
    t = np.zeros(n, np.float32)
    L.jxlo_dequant_table(0, t.ctypes.data)
    assert n == 192 and np.all(t > 0)
    assert abs(t[0] - 1 / 3150.0) < 1e-9 and abs(t[64] - 1 / 560.0) < 1e-9 and abs(t[128] - 1 / 512.0) < 1e-9   # A.8 DCT8 band 0
    for table in range(1, 17):
        if table == 10:
            continue   # AFV: not restated
        m = L.jxlo_dequant_table(table, None)
        tt = np.zeros(m, np.float32)
        L.jxlo_dequant_table(table, tt.ctypes.data)
        assert np.all(np.isfinite(tt)) and np.all(tt > 0)


@pytest.mark.parametrize("kw", [dict(effort=3), dict(effort=7), dict(effort=5, use_prefix=1), dict(effort=7, distance=6.0), dict(lossless=1), dict(lossless=1, use_prefix=1)])
def test_encode_decode_identities(oracle, kw):
    img = oracle.synthetic_image(141, 99, seed=9, channels=4)
    d = oracle.decode(oracle.encode(img, **kw))
    assert d.pixels.shape == img.shape and np.array_equal(d.pixels[..., 3], img[..., 3])
    if kw.get("lossless"):
        assert np.array_equal(d.pixels, img)
    else:
        assert oracle.psnr(d.pixels[..., :3], img[..., :3]) > 27.0


def test_oracle_rejects_bad_input(oracle):
    with pytest.raises(oracle.OracleError):
        oracle.decode(b"\x00\x01\x02\x03")
    data = oracle.encode(oracle.synthetic_image(300, 200, seed=1), effort=3)
    with pytest.raises(oracle.OracleError):
        oracle.decode(data[: len(data) // 2])
    assert oracle.lib().jxlo_signature_check(data, len(data)) == 2
    assert oracle.lib().jxlo_signature_check(b"\xff\x0a", 2) == 1


def test_orientation_and_16bit_float_outputs(oracle):
    img = oracle.synthetic_image(40, 24, seed=3)
    for orient in range(1, 9):
        d = oracle.decode(oracle.encode(img, lossless=1, orientation=orient))
        assert (d.width, d.height) == ((24, 40) if orient >= 5 else (40, 24))
    base = oracle.decode(oracle.encode(img, lossless=1)).pixels
    assert np.array_equal(oracle.decode(oracle.encode(img, lossless=1, orientation=2)).pixels, base[:, ::-1])
    assert np.array_equal(oracle.decode(oracle.encode(img, lossless=1, orientation=3)).pixels, base[::-1, ::-1])
    assert np.array_equal(oracle.decode(oracle.encode(img, lossless=1, orientation=4)).pixels, base[::-1])
    assert np.array_equal(oracle.decode(oracle.encode(img, lossless=1, orientation=5)).pixels, base.transpose(1, 0, 2))
    f = img.astype(np.float32) / 255.0
    d16 = oracle.decode(oracle.encode(f, bits=16, effort=3))
    assert d16.sample_type == 1 and d16.pixels.dtype == np.uint16 and oracle.psnr(d16.pixels / 257.0, img) > 30
    dh = oracle.decode(oracle.encode(f, bits=16, exp_bits=5, effort=3))
    assert dh.sample_type == 2 and dh.pixels.dtype == np.float16
    df = oracle.decode(oracle.encode(f, bits=32, exp_bits=8, effort=3))
    assert df.sample_type == 3 and df.pixels.dtype == np.float32 and oracle.psnr(df.pixels * 255.0, img) > 30
