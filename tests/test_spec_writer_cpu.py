"""CPU tests: files hand-assembled by the independent bit-level writer (tests/jxl_spec_writer.py) must be read back by the
oracle (pixels) and by the product's host front-end (JxlB200PeekInfo: headers only, no GPU). This pins the container, header,
TOC, entropy-code and Modular syntax of both readers against a third, separately written statement of ISO/IEC 18181."""
import numpy as np
import pytest

import spec_cases


def expected_pixels(px):
    if isinstance(px, tuple) and px[0] == "premul16":   # reference output: straight alpha (N/Decoder/JxlDecoder.cpp:233), orientation applied
        a = px[1].astype(np.float64) / 65535.0
        rgb = a[..., :3] / np.maximum(a[..., 3:4], 1.0 / 67108864.0)
        out = np.concatenate([np.clip(rgb, 0, 1), a[..., 3:4]], axis=2) * 65535.0
        return np.rot90(out, -1), 1.0
    if isinstance(px, tuple) and px[0] == "within1":    # float blending rounded to 8 bits: a rounding flip at x.5 is allowed
        return px[1], 0.5
    return px, 0


CASES = spec_cases.cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reads_spec_writer_files(oracle, case):
    name, data, px, info = case
    d = oracle.decode(data)
    want, tol = expected_pixels(px)
    assert d.pixels.shape == want.shape, (d.pixels.shape, want.shape)
    err = np.abs(d.pixels.astype(np.float64) - want.astype(np.float64)).max()
    assert err <= tol + (0.5 if tol else 0), "%s: max error %g" % (name, err)
    if "name" in info:
        assert d.name == info["name"]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_host_front_end_reads_spec_writer_headers(pkg, case):
    name, data, px, info = case
    got = pkg.peek_info(data)
    for k, v in info.items():
        if k != "name":
            assert got[k] == v, (name, k, got[k], v)
    assert not got["is_container"]


def test_container_variants(oracle, pkg):
    for name, data, px, exif, xmps in spec_cases.containerised():
        d = oracle.decode(data)
        assert np.array_equal(d.pixels, px), name
        assert d.is_container and d.exif == exif and d.xmp == xmps, name
        assert pkg.peek_info(data)["is_container"]


def test_dc_only_vardct_frame_matches_published_constants(oracle):
    data, want = spec_cases.vardct_dc_case()
    d = oracle.decode(data)
    assert d.pixels.shape == want.shape + () or d.pixels.shape[:2] == want.shape[:2]
    err = np.abs(d.pixels.astype(np.float64) - want).max()
    assert err <= 0.5 + 1e-2, "DC-only VarDCT frame: max |decoded - expected| = %g" % err
