"""GPU tests: files hand-assembled by the independent bit-level writer (tests/jxl_spec_writer.py) decoded through the C ABI
LoadImage. Lossless Modular files must come back bit-exact; the DC-only VarDCT frame within half an LSB of the float64 value
computed from the published XYB constants. The same files are checked against the oracle on CPU (test_spec_writer_cpu.py)."""
import numpy as np
import pytest

import spec_cases
from test_spec_writer_cpu import expected_pixels

pytestmark = pytest.mark.gpu

CASES = spec_cases.cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_load_image_decodes_spec_writer_files(gpu, case):
    name, data, px, info = case
    image = gpu.DecoderImage()
    gpu.JpegXLNative.LoadImage(data, image)
    got = image.layer_data.interleaved
    want, tol = expected_pixels(px)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).max()
    assert err <= tol + (0.5 if tol else 0), "%s: max error %g" % (name, err)
    assert image.format == info["format"] and image.has_transparency == info.get("has_transparency", False)
    if "name" in info:
        assert image.layer_data.name == info["name"].decode() + "\0"   # length passed includes the NUL (Appendix C-3)


def test_container_variants_through_load_image(gpu):
    for name, data, px, exif, xmps in spec_cases.containerised():
        image = gpu.DecoderImage()
        gpu.JpegXLNative.LoadImage(data, image)
        assert np.array_equal(image.layer_data.interleaved, px), name
        assert image.exif == exif and image.xmp == xmps[0], name
        assert image.callback_log.count("setXmp") == len(xmps) and image.callback_log.count("setExif") == 1


def test_dc_only_vardct_frame_matches_published_constants(gpu):
    data, want = spec_cases.vardct_dc_case()
    image = gpu.DecoderImage()
    gpu.JpegXLNative.LoadImage(data, image)
    err = np.abs(image.layer_data.interleaved.astype(np.float64) - want).max()
    assert err <= 0.5 + 1e-2, "DC-only VarDCT frame: max |decoded - expected| = %g" % err
