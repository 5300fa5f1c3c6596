"""CPU tests of the drop-in boundary and the host-side mirror of the managed interop layer (no GPU compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "JxlFileTypeIO.h")).read()
    declared = re.findall(r"JXLFT_API\s+[\w\s\*]+?JXLFT_CALL\s+(\w+)\s*\(", header)
    assert set(declared) >= {"GetLibJxlVersion", "LoadImage", "SaveImage"} and len(declared) >= 10
    lib = C.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(pkg.EXPORTS)


def test_struct_sizes_match_the_reference_abi(pkg):
    # LP64 sizes verified in SURVEY.md §8b against N/Common.h:17-60, N/Decoder/JxlDecoderTypes.h:63-71, N/Encoder/JxlEncoderTypes.h:27-42
    assert C.sizeof(pkg.BitmapData) == 24 and C.sizeof(pkg.EncoderOptionsNative) == 12 and C.sizeof(pkg.EncoderImageMetadataNative) == 48
    assert C.sizeof(pkg.DecoderCallbacks) == 48 and C.sizeof(pkg.IOCallbacks) == 16 and C.sizeof(pkg.ErrorInfo) == 256
    assert pkg.DECODER_STATUS.index("InvalidFileSignature") == 12 and pkg.DECODER_STATUS.index("DecodeError") == 10
    assert pkg.ENCODER_STATUS == ["Ok", "NullParameter", "OutOfMemory", "UserCancelled", "EncodeError", "WriteError"]
    assert pkg.KNOWN_COLOR_PROFILE.index("Rec2020PQ") == 7


def test_version_is_packed_like_jpegxl_numeric_version(pkg):
    major, minor, patch = pkg.JpegXLNative.GetLibJxlVersion()
    assert (major, minor) == (0, 11)


def test_null_parameters_and_signature_gate(pkg, oracle):
    lib = C.CDLL(pkg.LIB_PATH)
    lib.LoadImage.restype = C.c_int32
    lib.SaveImage.restype = C.c_int32
    ei = pkg.ErrorInfo()
    assert lib.LoadImage(None, None, C.c_size_t(0), C.byref(ei)) == 1          # NullParameter (N/Decoder/JxlDecoder.cpp:802-805)
    assert lib.SaveImage(None, None, None, None, C.byref(ei), None) == 1       # NullParameter (N/Encoder/JxlEncoder.cpp:155-158)
    image = pkg.DecoderImage()
    with pytest.raises(pkg.FormatException) as e:
        pkg.JpegXLNative.LoadImage(b"definitely not a jxl", image)
    assert e.value.status == "InvalidFileSignature" and image.callback_log == []


def test_peek_info_format_decisions(pkg, oracle):
    cases = [(dict(channels=3), dict(effort=3), ("Rgb", 0, False, 3)), (dict(channels=4), dict(lossless=1), ("Rgb", 0, True, 4)),
             (dict(channels=1), dict(effort=3), ("Gray", 0, False, 1))]
    for img_kw, enc_kw, want in cases:
        info = pkg.peek_info(oracle.encode(oracle.synthetic_image(40, 32, seed=1, **img_kw), **enc_kw))
        assert (info["format"], info["representation"], info["has_transparency"], info["num_channels"]) == want
        assert info["known_profile"] in ("Srgb", "GraySrgbTRC") and info["is_container"]
    f = oracle.synthetic_image(40, 32, seed=1).astype(np.float32) / 255
    assert pkg.peek_info(oracle.encode(f, bits=16, effort=3))["representation"] == 1          # Uint16
    assert pkg.peek_info(oracle.encode(f, bits=16, exp_bits=5, effort=3))["representation"] == 2   # Float16
    assert pkg.peek_info(oracle.encode(f, bits=32, exp_bits=8, effort=3))["representation"] == 3   # Float32
    with pytest.raises(pkg.FormatException) as e:
        pkg.peek_info(oracle.encode(f, bits=24, effort=3))
    assert "Unsupported integer bit depth: 24." in str(e.value)                                # N/Decoder/JxlDecoder.cpp:552
    p3 = pkg.peek_info(oracle.encode(f, bits=16, effort=3, primaries=11))
    assert p3["known_profile"] == "DisplayP3"
    pq = pkg.peek_info(oracle.encode(f, bits=16, effort=3, primaries=9, tf=16))
    assert pq["known_profile"] == "Rec2020PQ"
    swapped = pkg.peek_info(oracle.encode(oracle.synthetic_image(40, 32, seed=1), lossless=1, orientation=6))
    assert (swapped["width"], swapped["height"]) == (32, 40)


def test_without_a_gpu_the_engine_fails_loudly(pkg, oracle):
    ok, why = pkg.cuda_available()
    if ok:
        pytest.skip("a CUDA device is present")
    data = oracle.encode(oracle.synthetic_image(40, 32, seed=1), effort=3)
    image = pkg.DecoderImage()
    with pytest.raises(pkg.FormatException) as e:
        pkg.JpegXLNative.LoadImage(data, image)
    assert e.value.status == "DecodeError" and "no CPU fallback" in str(e.value)
    assert image.callback_log[:2] == ["setBasicInfo", "setKnownColorProfile"]   # pass 1 ran, pass 2 could not
    with pytest.raises(pkg.FormatException) as e:
        pkg.encode_to_memory(np.zeros((8, 8, 4), np.uint8), pkg.EncoderOptions())
    assert e.value.status == "EncodeError" and "no CPU fallback" in str(e.value)


def test_quality_to_distance_table(pkg):
    # I/QualityToDistanceLookupTable.cs:26-65, spot values from SURVEY.md §8c
    want = {100: 0.1, 95: 0.55, 90: 1.0, 75: 2.35, 50: 4.6, 30: 6.4, 29: 6.5922, 20: 7.4, 9: 13.907, 8: 15.0, 0: 15.0}
    for q, d in want.items():
        assert abs(pkg.quality_to_distance(q) - d) < 2e-3, q
    assert pkg.EncoderOptions(quality=10, lossless=True).distance == 0.0


def test_transparency_mapping(pkg):
    # I/TransparencyMapping.cs:19-32,43-55
    assert list(pkg.alpha_to_eight_bit(np.array([65535, 257, 256, 0], np.uint16))) == [255, 1, 0, 0]
    assert list(pkg.alpha_to_eight_bit(np.array([0.999, 1.0, 2.0, -1.0, 0.5], np.float32))) == [254, 255, 255, 0, 127]
    assert list(pkg.alpha_to_eight_bit(np.array([1.0, 0.5, 0.0], np.float16))) == [255, 127, 0]


def test_strip_heights(pkg):
    # S/BitmapUtil2.cs:61-71: max(1, 256 MiB / stride); SURVEY.md §8a row a9
    rects = list(pkg.enumerate_lock_rects(3840, 50000, 24))
    assert rects[0] == (0, 0, 3840, 23301) and rects[-1][3] == 50000
    assert list(pkg.enumerate_lock_rects(32768, 3000, 24))[0][3] == 2730
    assert list(pkg.enumerate_lock_rects(10, 5, 24)) == [(0, 0, 10, 5)]


def test_decoder_layer_data_split(pkg):
    rgba16 = np.arange(2 * 3 * 4, dtype=np.uint16).reshape(2, 3, 4) * 2570
    layer = pkg.DecoderLayerData(rgba16.tobytes(), "n\0", 3, 2, "Rgb", 1, True)
    assert layer.color.shape == (2, 3, 3) and layer.color.dtype == np.uint16
    assert np.array_equal(layer.transparency, (rgba16[..., 3] // 257).astype(np.uint8))
    gray = pkg.DecoderLayerData(bytes(range(6)), None, 3, 2, "Gray", 0, False)
    assert gray.color.shape == (2, 3, 3) and np.array_equal(gray.color[..., 0], gray.color[..., 2])


def test_shard_indices(pkg):
    files = list(range(10))
    shards = [pkg.shard_indices(len(files), r, 4) for r in range(4)]
    assert sorted(i for s in shards for i in s) == files and shards[1] == [1, 5, 9]


def test_colour_profile_callbacks_without_gpu(pkg):
    """Pass 1 of LoadImage (headers + metadata callbacks, N/Decoder/JxlDecoder.cpp:412-793) is host code: even without a CUDA device the
    profile callbacks fire before the decode fails. setKnownColorProfile xor setIccProfile; embedded ICC streams come back byte for
    byte; enum encodings outside the 8 known profiles get a synthesised matrix/TRC profile."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import oracle_py as O
    import icc_util
    ok, _ = pkg.cuda_available()
    img = O.synthetic_image(64, 48, seed=1)

    def load(data):
        im = pkg.DecoderImage()
        try:
            pkg.JpegXLNative.LoadImage(data, im)
        except pkg.FormatException:
            assert not ok          # only acceptable when there is no GPU: the engine has no CPU fallback
        return im

    icc = icc_util.make_matrix_icc(icc_util.ADOBE_PRIMS, gamma=2.19921875, curve="curv1")
    im = load(O.encode(img, lossless=True, icc=icc))
    assert (im.width, im.height) == (64, 48) and im.icc_profile == icc and im.known_color_profile is None
    im = load(O.encode(img, lossless=True))
    assert im.known_color_profile == "Srgb" and im.icc_profile is None
    im = load(O.encode(img, lossless=True, primaries=11, tf=1))           # P3 primaries + Rec.709 curve: not one of the 8 enums
    assert im.known_color_profile is None and im.icc_profile is not None
    p = icc_util.parse_icc(im.icc_profile)
    assert p["size"] == len(im.icc_profile) and p["space"] == b"RGB " and p["pcs"] == b"XYZ " and p["cls"] == b"mntr"
    want = icc_util.adapt_to_d50(0.3127, 0.3290) @ icc_util.rgb_to_xyz(icc_util.P3_PRIMS, 0.3127, 0.3290)
    got = np.stack([icc_util.xyz_of(p["tags"][s]) for s in (b"rXYZ", b"gXYZ", b"bXYZ")], axis=1)
    assert np.abs(got - want).max() < 2e-4
    t, params = icc_util.para_of(p["tags"][b"rTRC"])
    assert t == 3 and np.allclose(params, [1 / 0.45, 1 / 1.099, 0.099 / 1.099, 1 / 4.5, 0.081], atol=2e-4)
    im = load(O.encode(img, lossless=True, primaries=9, tf=16))          # Rec.2020 PQ is a known enum
    assert im.known_color_profile == "Rec2020PQ"


def test_band_layout_and_partition_are_host_only(pkg):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_py as O
    data = O.encode(O.synthetic_image(520, 700, seed=2), effort=3)
    assert pkg.band_layout(data) == (520, 700, 256, 3)
    data = O.encode(O.synthetic_image(300, 300, seed=3, channels=4), lossless=True)
    w, h, gdim, rows = pkg.band_layout(data)
    assert (w, h) == (300, 300) and gdim in (128, 256, 512, 1024) and rows == -(-300 // gdim)
    assert pkg.band_partition(128, 8) == [(16 * r, 16 * r + 16) for r in range(8)]      # BASELINE config 5: rows [16r, 16r+16) on GPU r
    assert pkg.band_partition(5, 2) == [(0, 3), (3, 5)] and pkg.band_partition(1, 3) == [(0, 1), (0, 0), (0, 0)]


def test_matrix_trc_icc_reading_for_lossy_encode(pkg):
    """What SaveImage derives from an ICC source profile when it has to reach XYB without a CMS (host code, no GPU needed): the colorant
    matrix to linear sRGB and the three tone curves, for 'para' types 0 and 3, a one-entry 'curv' gamma and a sampled 'curv' table."""
    import struct, sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import icc_util
    v = np.arange(256) / 255.0
    to_srgb = lambda prims: np.linalg.inv(icc_util.rgb_to_xyz(icc_util.SRGB_PRIMS, 0.3127, 0.3290)) @ icc_util.rgb_to_xyz(prims, 0.3127, 0.3290)
    for curve, gamma in (("para", 2.2), ("curv1", 2.19921875), ("table", 2.4)):
        m, lut = pkg.debug_parse_icc(icc_util.make_matrix_icc(icc_util.P3_PRIMS, gamma=gamma, curve=curve))
        assert np.abs(m - to_srgb(icc_util.P3_PRIMS)).max() < 2e-3            # s15Fixed16 colorants + Bradford round trip
        assert np.abs(lut - (v ** gamma)[None, :]).max() < (2e-4 if curve != "table" else 2e-3)
    # an sRGB profile with the parametric sRGB curve (type 3) must come out as the identity matrix and the sRGB decoding curve
    icc = bytearray(icc_util.make_matrix_icc(icc_util.SRGB_PRIMS, gamma=2.4, curve="para"))
    p = icc_util.parse_icc(bytes(icc))
    para3 = b"para" + b"\0" * 4 + struct.pack(">HH", 3, 0) + b"".join(icc_util.s15(x) for x in (2.4, 1 / 1.055, 0.055 / 1.055, 1 / 12.92, 0.04045))
    # rebuild the profile with the longer curve element appended at the end and the three TRC tags pointing at it
    n = struct.unpack(">I", icc[128:132])[0]
    off = len(icc)
    icc += para3
    for i in range(n):
        e = 132 + 12 * i
        if icc[e + 1: e + 4] == b"TRC":
            icc[e + 4: e + 12] = struct.pack(">II", off, len(para3))
    icc[0:4] = struct.pack(">I", len(icc))
    m, lut = pkg.debug_parse_icc(bytes(icc))
    assert np.abs(m - np.eye(3)).max() < 2e-3
    srgb = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
    assert np.abs(lut - srgb[None, :]).max() < 3e-4
    # profiles the engine must refuse for lossy encoding
    bad = bytearray(icc_util.make_matrix_icc(icc_util.P3_PRIMS)); bad[16:20] = b"CMYK"
    with pytest.raises(pkg.FormatException):
        pkg.debug_parse_icc(bytes(bad))
    bad = bytearray(icc_util.make_matrix_icc(icc_util.P3_PRIMS)); bad[20:24] = b"Lab "
    with pytest.raises(pkg.FormatException):
        pkg.debug_parse_icc(bytes(bad))
    with pytest.raises(pkg.FormatException):
        pkg.debug_parse_icc(b"\0" * 64)


def test_synthesised_profile_reads_back(pkg):
    """The profile the decoder synthesises for an enum encoding (P3 primaries, Rec.709 curve) fed to the encoder-side reader gives back the
    P3 -> sRGB matrix and the 709 decoding curve: the two host-side halves of the colour-profile handling agree with each other."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    import oracle_py as O
    import icc_util
    im = pkg.DecoderImage()
    try:
        pkg.JpegXLNative.LoadImage(O.encode(O.synthetic_image(32, 32, seed=5), lossless=True, primaries=11, tf=1), im)
    except pkg.FormatException:
        assert not pkg.cuda_available()[0]
    m, lut = pkg.debug_parse_icc(im.icc_profile)
    want = np.linalg.inv(icc_util.rgb_to_xyz(icc_util.SRGB_PRIMS, 0.3127, 0.3290)) @ icc_util.rgb_to_xyz(icc_util.P3_PRIMS, 0.3127, 0.3290)
    assert np.abs(m - want).max() < 2e-3
    v = np.arange(256) / 255.0
    bt709 = np.where(v < 0.081, v / 4.5, ((v + 0.099) / 1.099) ** (1 / 0.45))
    assert np.abs(lut - bt709[None, :]).max() < 3e-4


def test_round2_entry_points_without_a_gpu(pkg, oracle):
    """The entry points added in round 2 keep the boundary's rules: plain pointers, a status for every failure, a message, never a CPU
    fallback. (On a GPU box this test only checks the host-only parts.)"""
    import ctypes as C
    import numpy as np
    L = pkg._lib
    ok, _ = pkg.cuda_available()
    data = oracle.encode(oracle.synthetic_image(40, 30, seed=1, channels=4), lossless=1)
    # JxlB200LoadImageLayers: null parameters first, then the device requirement
    ei = pkg.ErrorInfo()
    assert L.JxlB200LoadImageLayers(None, 0, None, 0, None, 0, None, C.byref(ei)) == pkg.DECODER_STATUS.index("NullParameter")
    if not ok:
        with pytest.raises(pkg.FormatException) as e:
            pkg.load_image_layers(data)
        assert e.value.status == "DecodeError" and "CUDA" in str(e.value)
        # a band encoder session cannot be created without a device: NULL + message, no CPU path
        with pytest.raises(pkg.FormatException) as e:
            pkg.BandEncoder(np.zeros((16, 16, 4), np.uint8), 16, 0, 0, 0, pkg.EncoderOptions(quality=90, effort=3))
        assert "CUDA" in str(e.value)
    # JxlB200AssembleBands is host-only: garbage blobs and a histogram of the wrong size are refused with a message
    opts = pkg.EncoderOptions(quality=90, effort=3)
    with pytest.raises(pkg.FormatException) as e:
        pkg.assemble_bands(64, 64, opts, 0, np.zeros(8, np.uint64), [b"not a section blob"])
    assert e.value.status == "EncodeError"
    # section sizes of a file (host-only instrumentation): LfGlobal .. groups in logical order
    sizes, nlf, ng = pkg.section_sizes(oracle.encode(oracle.synthetic_image(300, 200, seed=2), effort=3))
    assert (nlf, ng) == (1, 2) and len(sizes) == 2 + nlf + ng and all(v > 0 for v in sizes)
    # the band partition of the sharded encoder: whole LF-group rows, halos of 8 rows towards every neighbour
    assert pkg.encode_band_partition(5000, 4) == [(0, 2048), (2048, 2048), (4096, 904), (0, 0)]
    assert pkg.band_rows_with_halo(5000, 2048, 2048) == (2040, 8, 8) and pkg.band_rows_with_halo(5000, 4096, 904) == (4088, 8, 0)
