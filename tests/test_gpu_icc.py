"""-m gpu tests of colour-profile handling at the boundary: ICC streams both ways, ICC synthesis for
enumerated encodings outside the 8 KnownColorProfile values, and lossy encoding from a matrix/TRC ICC source without a CMS."""
import io

import numpy as np
import pytest

import icc_util

pytestmark = pytest.mark.gpu


def _bgra(img):
    h, w, _ = img.shape
    out = np.empty((h, w, 4), np.uint8)
    out[..., 0], out[..., 1], out[..., 2], out[..., 3] = img[..., 2], img[..., 1], img[..., 0], 255
    return out


def _load(P, data):
    image = P.DecoderImage()
    P.JpegXLNative.LoadImage(data, image)
    return image


def test_lossless_icc_round_trip(gpu, oracle):
    icc = icc_util.make_matrix_icc(icc_util.ADOBE_PRIMS, gamma=2.19921875, curve="curv1")
    img = oracle.synthetic_image(333, 222, seed=21)
    out = io.BytesIO()
    gpu.JpegXLSave.Save(_bgra(img), out, lossless=True, icc=icc)
    data = out.getvalue()
    image = _load(gpu, data)
    assert image.icc_profile == icc and image.known_color_profile is None          # setIccProfile xor setKnownColorProfile
    assert np.array_equal(image.layer_data.color, img)
    d = oracle.decode(data)                                                         # the oracle reads the same ICC stream
    assert d.icc == icc and np.array_equal(d.pixels, img)
    # a gray-looking image with an ICC profile stays RGB (N/Encoder/JxlEncoder.cpp:67)
    gray = np.repeat(oracle.synthetic_image(64, 48, seed=2, channels=1), 3, axis=2)
    out = io.BytesIO()
    gpu.JpegXLSave.Save(_bgra(gray), out, lossless=True, icc=icc)
    assert _load(gpu, out.getvalue()).layer_data.color.shape[2] == 3


def test_oracle_written_icc_stream_is_reported(gpu, oracle):
    icc = icc_util.make_matrix_icc(icc_util.P3_PRIMS, gamma=2.2, curve="table")      # 2 KB sampled curve: exercises the entropy coder
    img = oracle.synthetic_image(200, 150, seed=22)
    image = _load(gpu, oracle.encode(img, lossless=True, icc=icc))
    assert image.icc_profile == icc and np.array_equal(image.layer_data.color, img)


def test_synthesised_icc_for_enum_encodings(gpu, oracle):
    """P3 primaries with the Rec.709 curve is expressible by the codestream enums but is none of the 8 known profiles: the engine
    reports a synthesised matrix/TRC profile, as libjxl does for the reference (N/Decoder/JxlDecoder.cpp:600-631)."""
    img = oracle.synthetic_image(160, 120, seed=23)
    data = oracle.encode(img, lossless=True, primaries=11, tf=1)
    image = _load(gpu, data)
    assert image.known_color_profile is None and image.icc_profile is not None
    p = icc_util.parse_icc(image.icc_profile)
    assert p["size"] == len(image.icc_profile) and p["version"] == 4 and p["cls"] == b"mntr" and p["space"] == b"RGB " and p["pcs"] == b"XYZ "
    want = icc_util.adapt_to_d50(0.3127, 0.3290) @ icc_util.rgb_to_xyz(icc_util.P3_PRIMS, 0.3127, 0.3290)
    got = np.stack([icc_util.xyz_of(p["tags"][s]) for s in (b"rXYZ", b"gXYZ", b"bXYZ")], axis=1)
    assert np.abs(got - want).max() < 2e-4
    assert np.abs(got.sum(axis=1) - np.array(icc_util.D50)).max() < 3e-4            # colorants sum to the PCS white
    t, params = icc_util.para_of(p["tags"][b"rTRC"])
    assert t == 3 and np.allclose(params, [1 / 0.45, 1 / 1.099, 0.099 / 1.099, 1 / 4.5, 0.081], atol=2e-4)
    assert p["tags"][b"gTRC"] == p["tags"][b"rTRC"] == p["tags"][b"bTRC"]
    assert np.array_equal(image.layer_data.color, img)                              # samples stay in the original encoding
    # known enums still go through setKnownColorProfile
    image = _load(gpu, oracle.encode(img, lossless=True, primaries=11, tf=13))
    assert image.known_color_profile == "DisplayP3" and image.icc_profile is None
    # gray with the 709 curve: kTRC profile
    g = oracle.synthetic_image(80, 60, seed=24, channels=1)
    image = _load(gpu, oracle.encode(g, lossless=True, color_space=1, tf=1))
    p = icc_util.parse_icc(image.icc_profile)
    assert p["space"] == b"GRAY" and b"kTRC" in p["tags"]


@pytest.mark.parametrize("curve", ["para", "curv1", "table"])
def test_lossy_encode_from_matrix_icc_source(gpu, oracle, curve):
    """SaveImage with a Display-P3-primaries, gamma-2.2 profile: the engine maps the samples to XYB from the profile's colorants and
    curves (libjxl uses a CMS here). Decoding gives sRGB samples (Appendix C-1); compare with the colorimetric conversion done in numpy."""
    icc = icc_util.make_matrix_icc(icc_util.P3_PRIMS, gamma=2.2, curve=curve)
    img = (oracle.synthetic_image(400, 300, seed=25).astype(np.float64) * 0.8 + 25).astype(np.uint8)      # keep clear of the gamut edge
    out = io.BytesIO()
    gpu.JpegXLSave.Save(_bgra(img), out, quality=95, effort=7, icc=icc)
    data = out.getvalue()
    image = _load(gpu, data)
    assert image.known_color_profile == "Srgb" and image.icc_profile is None              # sRGB samples are reported as sRGB
    lin = (img / 255.0) ** 2.2
    m = np.linalg.inv(icc_util.rgb_to_xyz(icc_util.SRGB_PRIMS, 0.3127, 0.3290)) @ icc_util.rgb_to_xyz(icc_util.P3_PRIMS, 0.3127, 0.3290)
    s = np.clip(lin @ m.T, 0, 1)
    want = np.round(255 * np.where(s <= 0.0031308, 12.92 * s, 1.055 * s ** (1 / 2.4) - 0.055))
    got = image.layer_data.color.astype(np.float64)
    assert oracle.psnr(got, want) >= 38.0
    ref = oracle.decode(data, threads=4).pixels
    assert int(np.abs(ref.astype(np.int32) - image.layer_data.color.astype(np.int32)).max()) <= 1
    # the profile itself is in the file for decoders that do have a CMS
    assert oracle.decode(data).icc == icc


def test_lossy_encode_refuses_profiles_it_cannot_read(gpu, oracle):
    icc = bytearray(icc_util.make_matrix_icc(icc_util.P3_PRIMS))
    icc[20:24] = b"Lab "                                                             # LUT-based profile: needs a CMS
    img = oracle.synthetic_image(64, 48, seed=26)
    with pytest.raises(gpu.FormatException) as e:
        gpu.JpegXLSave.Save(_bgra(img), io.BytesIO(), quality=90, icc=bytes(icc))
    assert "CMS" in str(e.value)
    out = io.BytesIO()                                                               # ... while lossless just carries it
    gpu.JpegXLSave.Save(_bgra(img), out, lossless=True, icc=bytes(icc))
    assert _load(gpu, out.getvalue()).icc_profile == bytes(icc)
