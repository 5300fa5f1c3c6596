"""The one reference translation unit that compiles without libjxl (N/Encoder/PixelFormatConversion.cpp:16-121, built in place by
oracle/Makefile into oracle/_ref/) against a numpy restatement — the bit-exact oracle for the encoder's BGRA repack."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_py


class BitmapData(C.Structure):
    _fields_ = [("scan0", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("stride", C.c_uint32)]


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(oracle_py.REF_LIB_PATH):
        oracle_py.build()
    if not os.path.exists(oracle_py.REF_LIB_PATH):
        pytest.skip("oracle/_ref not built (reference sources absent on this box)")
    return C.CDLL(oracle_py.REF_LIB_PATH)


@pytest.mark.parametrize("w,h,pad", [(7, 5, 0), (33, 9, 12), (1, 1, 4)])
def test_bgra_repacks_match_numpy_restatement(ref, w, h, pad):
    rng = np.random.default_rng(w * h)
    stride = w * 4 + pad
    buf = rng.integers(0, 256, (h, stride), dtype=np.uint8)
    px = buf[:, : w * 4].reshape(h, w, 4)
    bm = BitmapData(buf.ctypes.data, w, h, stride)
    for fn, nch, want in (("ref_BgraToGray", 1, px[..., 0:1]), ("ref_BgraToGrayAlpha", 2, px[..., [0, 3]]),
                          ("ref_BgraToRgb", 3, px[..., [2, 1, 0]]), ("ref_BgraToRgba", 4, px[..., [2, 1, 0, 3]])):
        out = np.zeros((h, w, nch), np.uint8)
        getattr(ref, fn)(C.byref(bm), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, want), fn
