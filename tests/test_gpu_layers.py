"""-m gpu parity tests for layered (multi-frame) stills: the reference receives the first full image of a coalescing decoder
(N/Decoder/JxlDecoder.cpp:252-400), i.e. every zero-duration frame blended onto the canvas / a reference slot until the first frame that
is shown. Files come from the oracle encoder's layer options; the GPU result is compared with the oracle decoder and, for the blend
formulas, with a composite computed here in numpy from the source layers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 200, 150


def _decode_gpu(P, data):
    image = P.DecoderImage()
    P.JpegXLNative.LoadImage(data, image)
    return image


from layer_util import composite as _composite, to_u8 as _to_u8  # noqa: E402


@pytest.mark.parametrize("mode", ["replace", "add", "blend", "muladd", "mul"])
@pytest.mark.parametrize("x0,y0", [(30, 20), (-20, 100), (150, -10)])
def test_lossless_layers_match_the_oracle_and_the_formulas(gpu, oracle, mode, x0, y0):
    base = oracle.synthetic_image(W, H, seed=1, channels=4)
    base[..., 3] = np.maximum(base[..., 3], 100)
    over = oracle.synthetic_image(90, 70, seed=2, channels=4)
    data = oracle.encode_layers(W, H, [(base, dict()), (over, dict(x0=x0, y0=y0, mode=mode))], lossless=1)
    ref = oracle.decode(data).pixels
    image = _decode_gpu(gpu, data)
    got = image.layer_data.interleaved
    assert got.shape == ref.shape == (H, W, 4) and image.has_transparency
    assert int(np.abs(got.astype(np.int32) - ref.astype(np.int32)).max()) <= 1
    want = _to_u8(_composite(base.astype(np.float32) / 255.0, over.astype(np.float32) / 255.0, x0, y0, mode))
    assert int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max()) <= 1
    if mode == "replace":
        assert np.array_equal(got, want)


@pytest.mark.parametrize("kw,tol", [(dict(effort=7), 1), (dict(effort=3), 1), (dict(effort=5, premultiplied=1), 1)])
def test_lossy_layers(gpu, oracle, kw, tol):
    """VarDCT frames as layers: each goes through the full pipeline (gaborish + EPF at effort >= 5) into float samples before it is blended."""
    base = oracle.synthetic_image(320, 280, seed=3, channels=4)
    base[..., 3] = np.maximum(base[..., 3], 60)
    over = oracle.synthetic_image(150, 130, seed=4, channels=4)
    if kw.get("premultiplied"):
        for im in (base, over):
            im[..., :3] = (im[..., :3].astype(np.float32) * im[..., 3:4] / 255.0).astype(np.uint8)
    data = oracle.encode_layers(320, 280, [(base, dict()), (over, dict(x0=100, y0=-30, mode="blend"))], **kw)
    ref = oracle.decode(data).pixels
    got = _decode_gpu(gpu, data).layer_data.interleaved
    assert got.shape == ref.shape
    assert int(np.abs(got.astype(np.int32) - ref.astype(np.int32)).max()) <= tol


def test_three_layers_through_reference_slots(gpu, oracle):
    """Layer 1 is saved to slot 1, layer 2 starts from the empty slot 2 and is saved there, the last layer blends onto slot 1: slot 2's
    content must not leak into the result."""
    a = oracle.synthetic_image(W, H, seed=5, channels=4)
    b = oracle.synthetic_image(W, H, seed=6, channels=4)
    c = oracle.synthetic_image(80, 60, seed=7, channels=4)
    layers = [(a, dict(save=1)), (b, dict(source=2, save=2, mode="add")), (c, dict(x0=60, y0=40, source=1, mode="blend"))]
    data = oracle.encode_layers(W, H, layers, lossless=1)
    ref = oracle.decode(data).pixels
    got = _decode_gpu(gpu, data).layer_data.interleaved
    assert int(np.abs(got.astype(np.int32) - ref.astype(np.int32)).max()) <= 1
    want = _to_u8(_composite(a.astype(np.float32) / 255.0, c.astype(np.float32) / 255.0, 60, 40, "blend"))
    assert int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max()) <= 1


@pytest.mark.parametrize("kw,dtype", [(dict(bits=16), np.uint16), (dict(bits=32, exp_bits=8), np.float32)])
def test_layers_of_gray_alpha_images_at_other_sample_types(gpu, oracle, kw, dtype):
    base = oracle.synthetic_image(W, H, seed=8, channels=4)[..., [0, 3]].astype(np.float32) / 255.0
    over = oracle.synthetic_image(70, 90, seed=9, channels=4)[..., [1, 3]].astype(np.float32) / 255.0
    data = oracle.encode_layers(W, H, [(base, dict()), (over, dict(x0=10, y0=30, mode="blend"))], lossless=1, num_color=1, has_alpha=True, orientation=6, **kw)
    ref = oracle.decode(data).pixels
    image = _decode_gpu(gpu, data)
    got = image.layer_data.interleaved
    assert got.dtype == dtype and got.shape == ref.shape == (W, H, 2) and image.format == "Gray"      # orientation 6: rotated
    if dtype == np.uint16:
        assert int(np.abs(got.astype(np.int64) - ref.astype(np.int64)).max()) <= 2
    else:
        assert float(np.abs(got - ref).max()) <= 1e-5


def test_layered_file_to_bgra_surface(gpu, oracle):
    base = oracle.synthetic_image(W, H, seed=10, channels=4)
    over = oracle.synthetic_image(64, 64, seed=11, channels=4)
    data = oracle.encode_layers(W, H, [(base, dict()), (over, dict(x0=5, y0=5, mode="blend"))], lossless=1)
    doc = gpu.JpegXLLoad.Load(data)
    fused = gpu.load_image_bgra(data)
    assert int(np.abs(doc.surface.astype(np.int32) - fused.astype(np.int32)).max()) <= 1


def test_layered_files_inside_a_batch_and_band_refusal(gpu, oracle):
    """A batch may contain layered stills: they are composited by the single-image path while the other files go through the phased pipeline.
    Band decode (one frame cut by group rows) has no meaning for a layered file and says so."""
    base = oracle.synthetic_image(W, H, seed=12, channels=4)
    layered = oracle.encode_layers(W, H, [(base, dict()), (base[:50, :50], dict(x0=20, y0=30, mode="add"))], lossless=1)
    plain = oracle.encode(base, lossless=1)
    files = [plain, layered, plain, layered]
    outs = [np.zeros((H, W, 4), np.uint8) for _ in files]
    st = gpu.decode_batch(files, outs)
    assert list(st) == [0, 0, 0, 0]
    want_layered = oracle.decode(layered).pixels
    assert np.array_equal(outs[0], base) and np.array_equal(outs[2], base)
    assert int(np.abs(outs[1].astype(np.int32) - want_layered.astype(np.int32)).max()) <= 1 and np.array_equal(outs[1], outs[3])
    with pytest.raises(gpu.FormatException, match="layered"):
        gpu.band_layout(layered)
