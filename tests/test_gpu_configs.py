"""-m gpu tests at the full sizes of BASELINE.json's configs 2 and 4 (config 1 / 3 are the small-size and batch tests of
test_gpu_decode.py / bench.py; config 5, the gigapixel frame sharded by group rows, is covered at test size by
test_band_decode_stitches_to_the_full_frame and at full size by scripts/gigapixel_bands.py — DESIGN.md §7)."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bgra(img):
    h, w, c = img.shape
    out = np.empty((h, w, 4), np.uint8)
    out[..., 0], out[..., 1], out[..., 2] = img[..., 2], img[..., 1], img[..., 0]
    out[..., 3] = img[..., 3] if c == 4 else 255
    return out


def test_config2_uhd_decode_with_bgra_pack(gpu, oracle):
    """3840x2160 RGB8 VarDCT d=1.0 e=7 -> decode + BGRA32 surface on the GPU (the a7-a10 repack fused into the output kernel)."""
    w, h = 3840, 2160
    img = oracle.synthetic_image(w, h, seed=0)
    data = gpu.encode_to_memory(_bgra(img), gpu.EncoderOptions(quality=90, effort=7))
    surf = gpu.load_image_bgra(data)
    assert surf.shape == (h, w, 4) and int(surf[..., 3].min()) == 255
    ref = oracle.decode(data, threads=16).pixels
    assert int(np.abs(surf[..., 2::-1].astype(np.int32) - ref.astype(np.int32)).max()) <= 1      # <= 1 LSB against the CPU oracle
    assert oracle.psnr(surf[..., 2::-1], ref) >= 60.0
    assert oracle.psnr(ref, img) >= 36.0                                                         # and the file is a faithful d=1.0 encode
    image = gpu.DecoderImage()                                                                   # same pixels through LoadImage + the managed repack
    gpu.JpegXLNative.LoadImage(data, image)
    assert np.array_equal(image.layer_data.color, surf[..., 2::-1])


def test_config4a_lossless_rgba_4096(gpu, oracle):
    """4096x4096 RGBA8 lossless Modular; alpha = radial gradient x noise with fully transparent and fully opaque regions."""
    n = 4096
    rgb = oracle.synthetic_image(n, n, seed=0)
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32)
    radial = 3.0 * (1.28 - np.hypot(xx - n / 2, yy - n / 2) / (n / 2))
    alpha = np.clip(radial * (0.75 + 0.5 * rng.random((n, n), dtype=np.float32)), 0, 1)
    alpha[alpha > 0.8] = 1.0
    alpha[alpha < 0.08] = 0.0
    a8 = np.round(alpha * 255).astype(np.uint8)
    assert 0.02 < float((a8 == 0).mean()) < 0.2 and float((a8 == 255).mean()) > 0.4
    rgba = np.concatenate([rgb, a8[..., None]], axis=2)
    out = io.BytesIO()
    gpu.JpegXLSave.Save(_bgra(rgba), out, lossless=True, effort=7)
    data = out.getvalue()
    image = gpu.DecoderImage()
    gpu.JpegXLNative.LoadImage(data, image)
    assert image.layer_data.transparency is not None and np.array_equal(image.layer_data.color, rgb) and np.array_equal(image.layer_data.transparency, a8)
    d = oracle.decode(data, threads=16)                                                          # bit-exact against the CPU oracle as well
    assert np.array_equal(d.pixels, rgba)
    assert len(data) < rgba.nbytes * 0.8


@pytest.mark.parametrize("kw,dtype", [(dict(bits=16), np.uint16), (dict(bits=32, exp_bits=8), np.float32)])
def test_config4b_8k_hdr_pq(gpu, oracle, kw, dtype):
    """7680x4320 Rec.2020-PQ VarDCT with gaborish and two EPF iterations, 16-bit and float32 output."""
    w, h = 7680, 4320
    img = oracle.synthetic_image(w, h, seed=0).astype(np.float32) / 255.0
    data = oracle.encode(img, effort=3, gab=1, epf=2, primaries=9, tf=16, intensity_target=1000.0, threads=16, **kw)
    image = gpu.DecoderImage()
    gpu.JpegXLNative.LoadImage(data, image)
    got = image.layer_data.color
    assert got.dtype == dtype and got.shape == (h, w, 3) and image.known_color_profile == "Rec2020PQ"
    ref = oracle.decode(data, threads=16).pixels
    if dtype == np.uint16:
        assert int(np.abs(got.astype(np.int64) - ref.astype(np.int64)).max()) <= 257             # <= 1 LSB at 8-bit precision
    else:
        # The PQ curve has unbounded slope at black (E ~ L^0.159): a 1e-7 difference in linear light near zero — fp32 summation order in
        # the IDCT — moves the code value by 1e-3, so a uniform relative bound cannot hold for every sample of a 100 M sample frame.
        # Bounds: 99.8 % of the samples within 1e-4 relative (north_star), every sample within 1e-2 absolute of the code range, and
        # every sample within 2e-4 relative + 3e-6 of full scale once both outputs are taken back to linear light (the opsin inverse
        # has coefficients of +11 / -9.9, so fp32 cancellation alone gives a few 1e-5 relative there).
        a, b = got.astype(np.float64), ref.astype(np.float64)
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
        assert float((rel <= 1e-4).mean()) >= 0.998 and float(np.abs(a - b).max()) <= 1e-2

        def pq_to_linear(e):
            m1, m2, c1, c2, c3 = 2610.0 / 16384, 2523.0 / 4096 * 128, 3424.0 / 4096, 2413.0 / 4096 * 32, 2392.0 / 4096 * 32
            p = np.abs(e) ** (1 / m2)
            return np.sign(e) * (np.maximum(p - c1, 0) / (c2 - c3 * p)) ** (1 / m1)
        la, lb = pq_to_linear(a[::3, ::3]), pq_to_linear(b[::3, ::3])
        assert bool(np.all(np.abs(la - lb) <= 2e-4 * np.abs(lb) + 3e-6))
