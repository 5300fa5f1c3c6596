"""CPU tests with the two ICC profiles the reference ships (S/ColorProfiles/*.icc, used by I/DecoderImage.cs:166-184) as real-world vectors:
the oracle's ICC-stream codec must round-trip them byte for byte, and the product's host-side reader of matrix/TRC profiles (what a lossy
SaveImage with metadata->iccProfile uses, N/Encoder/JxlEncoder.cpp:258-262) must recover the published primaries and tone curves.
The profiles are read in place from /root/reference (never copied into the repo); the tests skip where the reference is absent."""
import glob
import os

import numpy as np
import pytest

PROFILE_DIR = "/root/reference/src/ColorProfiles"
PROFILES = sorted(glob.glob(os.path.join(PROFILE_DIR, "*.icc")))
pytestmark = pytest.mark.skipif(not PROFILES, reason="reference ICC profiles not present on this machine")


@pytest.mark.parametrize("path", PROFILES, ids=[os.path.basename(p) for p in PROFILES])
def test_icc_stream_codec_round_trips_the_reference_profiles(oracle, path):
    icc = open(path, "rb").read()
    assert icc[36:40] == b"acsp" and int.from_bytes(icc[0:4], "big") == len(icc)
    stream = oracle.icc_stream_write(icc)
    assert len(stream) < len(icc)                      # the predicted-ICC coding must actually compress a real profile
    assert oracle.icc_stream_read(stream) == icc


def test_matrix_trc_reader_on_the_reference_profiles(pkg):
    by_name = {os.path.basename(p): open(p, "rb").read() for p in PROFILES}
    # Rec.709 primaries and curve: same primaries as sRGB -> identity matrix to linear sRGB; the curve is the BT.709 inverse OETF
    m, lut = pkg.debug_parse_icc(by_name["Rec709-elle-V4-rec709.icc"])
    assert np.allclose(np.asarray(m).reshape(3, 3), np.eye(3), atol=2e-3)
    v = np.arange(256) / 255.0
    bt709 = np.where(v < 0.081, v / 4.5, ((v + 0.099) / 1.099) ** (1 / 0.45))
    assert np.allclose(np.asarray(lut).reshape(3, 256), bt709[None, :], atol=2e-3)
    # Rec.2020 primaries, linear curve: the published BT.2020 -> BT.709 matrix
    m, lut = pkg.debug_parse_icc(by_name["Rec2020-elle-V4-g10.icc"])
    want = np.array([[1.6605, -0.5876, -0.0728], [-0.1246, 1.1329, -0.0083], [-0.0182, -0.1006, 1.1187]])
    assert np.allclose(np.asarray(m).reshape(3, 3), want, atol=3e-3)
    assert np.allclose(np.asarray(lut).reshape(3, 256), v[None, :], atol=1e-4)
