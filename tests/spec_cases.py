"""Files assembled by the independent bit-level writer (tests/jxl_spec_writer.py) and the pixels each must decode to.

Every case exercises codestream syntax that the oracle's and the product's readers share text for (SizeHeader forms,
ImageMetadata fields, FrameHeader, TOC with and without a permutation, entropy-code headers, alias tables, prefix codes,
MA trees, RCT, container boxes, a DC-only VarDCT frame). Used by the CPU tests (oracle + host-only PeekInfo) and by the
GPU tests (LoadImage through the C ABI).
"""
import numpy as np

import jxl_spec_writer as W


def _planes(arr):
    """HxWxC array -> list of C channels, each a list of rows of Python ints."""
    a = np.asarray(arr)
    if a.ndim == 2:
        a = a[:, :, None]
    return [[[int(v) for v in row] for row in a[:, :, c]] for c in range(a.shape[2])]


def _img(h, w, c, seed, lo=0, hi=255, smooth=True):
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.integers(lo, hi + 1, size=(h, w, c)).astype(np.int64)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * (3 + k) + yy * (5 - k) + 17 * k) % (hi - lo + 1) for k in range(c)], axis=2)
    noise = rng.integers(-6, 7, size=(h, w, c))
    return np.clip(base + noise + lo, lo, hi).astype(np.int64)


def cases():
    """-> list of (name, file bytes, expected pixels (HxWxC, dtype of the reference's output), expected info dict)"""
    out = []

    # 1. gray 8-bit, "small" SizeHeader with an aspect ratio, single gradient leaf, flat 256-symbol ANS (trivial alias table)
    g = _img(24, 32, 1, seed=1)          # 32 = 24 * 4 / 3: ratio 3
    out.append(("gray8_small_ratio", W.modular_image(_planes(g), gray=True), g.astype(np.uint8), dict(width=32, height=24, format="Gray", num_channels=1)))

    # 2. RGB 8-bit, explicit U32 sizes (not a multiple of 8), frame name, no RCT, random (noisy) content
    a = _img(13, 21, 3, seed=2, smooth=False)
    out.append(("rgb8_u32_size_named", W.modular_image(_planes(a), name=b"layer one"), a.astype(np.uint8), dict(width=21, height=13, format="Rgb", num_channels=3, name=b"layer one")))

    # 3. RGB + 8-bit alpha (default ExtraChannelInfo), YCgCo RCT (type 6), tree splitting on property 0 (channel) and 9 (gradient)
    a = _img(16, 16, 4, seed=3)
    tree = W.Split(0, 2, W.Leaf(0, 1), W.Split(9, 40, W.Leaf(1, 5), W.Leaf(2, 4)))
    out.append(("rgba8_rct6_tree", W.modular_image(_planes(a), alpha_bits=8, tree=tree, rct=(0, 6)), a.astype(np.uint8), dict(width=16, height=16, format="Rgb", num_channels=4, has_transparency=True)))

    # 4. non-power-of-two flat ANS alphabet (needs the alias construction), 6-bit table, two-symbol and single-symbol clusters.
    #    Leaves are numbered in breadth-first order: ctx 0 = channel 0 (flat 40-symbol alphabet, hybrid config 5/1/1),
    #    ctx 1 = channel 2 (constant: single-symbol distribution), ctx 2 = channel 1 (two residual values: two-symbol distribution)
    h, w = 12, 20
    rng = np.random.default_rng(4)
    c0 = rng.integers(0, 64, size=(h, w))
    c1 = np.zeros((h, w), np.int64) + 77
    c1[:, 1::2] += 1
    c2 = np.zeros((h, w), np.int64) + 200
    a = np.stack([c0, c1, c2], axis=2)
    tree = W.Split(0, 0, W.Split(0, 1, W.Leaf(1, 0, offset=200), W.Leaf(2, 0, offset=77)), W.Leaf(0, 0))
    code = W.EntropyCode([0, 1, 2], [("flat", 40), ("single", 0), ("two", 0, 2, 1365)], hybrids=[W.Hybrid(5, 1, 1), W.Hybrid(4, 2, 0), W.Hybrid(0, 0, 0)], log_alpha=6)
    out.append(("rgb8_alias_tables", W.modular_image(_planes(a), tree=tree, data_code=code), a.astype(np.uint8), dict(width=w, height=h, format="Rgb", num_channels=3)))

    # 5. simple prefix codes (2, 3 and 4 symbols) and leaf multipliers: ctx 0 = channel 0 (4 levels x 60), ctx 1 = channel 2 (2 levels x 255),
    #    ctx 2 = channel 1 (3 levels x 100)
    rng = np.random.default_rng(5)
    a = np.stack([rng.integers(0, 4, size=(9, 14)) * 60, rng.integers(0, 3, size=(9, 14)) * 100, rng.integers(0, 2, size=(9, 14)) * 255], axis=2)
    tree = W.Split(0, 0, W.Split(0, 1, W.Leaf(1, 0, multiplier=255), W.Leaf(2, 0, multiplier=100)), W.Leaf(0, 0, multiplier=60))
    code = W.EntropyCode([0, 1, 2], [(0, 2, 4, 6), (0, 2), (0, 2, 4)], hybrids=[W.Hybrid(4, 1, 0)] * 3, use_prefix=True)
    out.append(("rgb8_prefix_codes", W.modular_image(_planes(a), tree=tree, data_code=code), a.astype(np.uint8), dict(width=14, height=9, format="Rgb", num_channels=3)))

    # 6. 16-bit RGB with 16-bit premultiplied alpha and orientation 6 (output is rotated); hybrid-uint tails in use
    a = _img(10, 18, 4, seed=6, hi=65535)
    a[..., 3] = np.maximum(a[..., 3], 30000)
    f = W.modular_image(_planes(a), bits=16, alpha_bits=16, orientation=6, alpha_associated=True,
                        data_code=W.EntropyCode([0], [("flat", 100)], hybrids=[W.Hybrid(4, 2, 0)], log_alpha=7))   # 100 symbols over 128 buckets: alias redirects
    out.append(("rgba16_orient6_premul", f, ("premul16", a), dict(width=10, height=18, format="Rgb", num_channels=4, representation=1, has_transparency=True)))

    # 7. multi-group frame (group size 128) with a real TOC, every group its own section
    a = _img(150, 200, 3, seed=7)
    out.append(("rgb8_multigroup_toc", W.modular_image(_planes(a), group_size_shift=0), a.astype(np.uint8), dict(width=200, height=150, format="Rgb", num_channels=3)))

    # 8. the same with the sections stored in a permuted order (TOC permutation, Lehmer-coded)
    nsec = 1 + 1 + 1 + 4
    order = [0, 1, 2, 6, 4, 3, 5]
    out.append(("rgb8_permuted_toc", W.modular_image(_planes(a), group_size_shift=0, toc_permutation=order), a.astype(np.uint8), dict(width=200, height=150, format="Rgb", num_channels=3)))
    assert len(order) == nsec

    # 9. CMYK: black extra channel (+ alpha), 8-bit; the reference inverts C, M, Y, K (N/Decoder/JxlDecoder.cpp:159-215)
    a = _img(11, 15, 5, seed=9, smooth=False)
    ecs = [dict(type=W.EC_BLACK, bits=8, name=b"K")]
    f = W.modular_image(_planes(a), alpha_bits=8, extra=ecs)          # channel order in the file: R,G,B(=C,M,Y), alpha, black
    want = np.concatenate([255 - a[..., 0:3], 255 - a[..., 4:5], a[..., 3:4]], axis=2).astype(np.uint8)
    out.append(("cmyka8", f, want, dict(width=15, height=11, format="Cmyk", num_channels=4, has_transparency=True)))
    # 10. LZ77: long runs of equal residuals become (length, distance 1) copies; ANS with a 256-symbol alphabet so that length tokens
    #     (min_symbol 224 + ...) fit; the distance context is one extra entry of the context map. Single group and multi-group.
    for nm, (hh, ww), shift in (("rgb8_lz77_runs", (40, 64), 1), ("rgb8_lz77_multigroup", (150, 200), 0)):
        yy, xx = np.mgrid[0:hh, 0:ww]
        a = np.stack([(xx // 16) * 9 + (yy // 8) * 3, np.where((xx > ww // 3) & (yy > hh // 4), 200, 17), (yy // 5) * 7 % 256], axis=2).astype(np.int64) % 256
        code = W.EntropyCode([0, 1], [("flat", 256), ("flat", 32)], hybrids=[W.Hybrid(4, 2, 0), W.Hybrid(4, 1, 0)], log_alpha=8,
                             lz77=dict(min_symbol=224, min_length=3, len_hybrid=W.Hybrid(2, 0, 0)))
        out.append((nm, W.modular_image(_planes(a), data_code=code, rle=(3, 1), group_size_shift=shift), a.astype(np.uint8), dict(width=ww, height=hh, format="Rgb", num_channels=3)))
    # 11. MA-tree properties of previous channels: channel 1 splits on the value of channel 0 (property 17), channel 2 on channel 1's
    #     gradient residual (property 19) and on channel 0's absolute value (property 20 = second previous channel); no RCT, 2 groups
    a = _img(140, 130, 3, seed=12)
    a[..., 1] = (a[..., 0] * 3 // 4 + a[..., 1] // 8) % 256
    tree = W.Split(0, 0, W.Split(0, 1, W.Split(19, 2, W.Leaf(2, 5), W.Split(20, 90, W.Leaf(5, 4), W.Leaf(6, 1))), W.Split(17, 100, W.Leaf(3, 5), W.Leaf(4, 2))), W.Leaf(0, 5))
    # breadth-first leaf order: root; [gt Split(ch>1?), le Leaf0]; then [Split(19), Split(17)]; then [Leaf2?...]: renumber by construction below
    def renumber(root):
        k = 0
        for n in W.tree_nodes_bfs(root):
            if isinstance(n, W.Leaf):
                n.ctx = k
                k += 1
        return k
    nl = renumber(tree)
    code = W.EntropyCode(list(range(nl)), [("flat", 64)] * nl, hybrids=[W.Hybrid(4, 2, 0)] * nl, log_alpha=6) if nl <= 8 else None
    out.append(("rgb8_prev_channel_props", W.modular_image(_planes(a), tree=tree, data_code=code, group_size_shift=0), a.astype(np.uint8), dict(width=130, height=140, format="Rgb", num_channels=3)))
    # 12. Palette (H.6.3): 37 explicit colours + colours of both implicit cubes; the palette is meta channel 0, the index channel replaces R,G,B.
    #     Single group; then 3 x 2 groups of 128 px (palette in the global section, index channel in the group sections); then with an
    #     alpha channel behind the palettised colour channels and an RCT applied before the palette.
    rng = np.random.default_rng(13)
    explicit = [tuple(int(v) for v in rng.integers(0, 256, size=3)) for _ in range(37)]
    cube = [(32, 95, 159), (223, 223, 32), (0, 63, 255), (191, 127, 0), (255, 255, 255), (95, 95, 95)]      # 4x4x4 levels 32/95/159/223, 5x5x5 levels 0/63/127/191/255
    def paletted(hh, ww, seed):
        r = np.random.default_rng(seed)
        pick = r.integers(0, len(explicit) + len(cube), size=(hh, ww))
        pick = np.where(r.random((hh, ww)) < 0.7, np.roll(pick, 1, axis=1), pick)                           # some horizontal coherence
        table = np.array(explicit + cube, dtype=np.int64)
        return table[pick]
    a = paletted(90, 120, 1)
    tree = W.Split(0, 0, W.Leaf(0, 1), W.Leaf(1, 0))                                                        # ctx 0: indices (channel > 0, predictor W), ctx 1: the palette rows (predictor Zero)
    code = W.EntropyCode([0, 1], [("flat", 128), ("flat", 256)], hybrids=[W.Hybrid(4, 2, 0), W.Hybrid(4, 2, 0)], log_alpha=8)
    pal = dict(begin=0, num_c=3, colors=explicit)
    out.append(("rgb8_palette", W.modular_image(_planes(a), tree=tree, data_code=code, palette=pal), a.astype(np.uint8), dict(width=120, height=90, format="Rgb", num_channels=3)))
    a = paletted(200, 300, 2)
    out.append(("rgb8_palette_groups", W.modular_image(_planes(a), tree=tree, data_code=code, palette=pal, group_size_shift=0), a.astype(np.uint8), dict(width=300, height=200, format="Rgb", num_channels=3)))
    a = np.concatenate([paletted(70, 50, 3), _img(70, 50, 1, seed=14)], axis=2)
    tree3 = W.Split(0, 1, W.Leaf(0, 5), W.Split(0, 0, W.Leaf(1, 1), W.Leaf(2, 0)))                          # BFS: alpha (gradient), indices (W), palette (Zero)
    renumber(tree3)
    code3 = W.EntropyCode([0, 1, 2], [("flat", 256), ("flat", 128), ("flat", 256)], hybrids=[W.Hybrid(4, 2, 0)] * 3, log_alpha=8)
    out.append(("rgba8_palette_alpha", W.modular_image(_planes(a), alpha_bits=8, tree=tree3, data_code=code3, palette=pal), a.astype(np.uint8), dict(width=50, height=70, format="Rgb", num_channels=4, has_transparency=True)))
    # 13. Palette with delta entries: the first two entries are deltas added to the W prediction of the reconstructed channel (H.6.3): a
    #     horizontal ramp is "previous pixel + (1, 2, 3)" / "previous pixel + (-2, 0, 1)" after an explicit first column
    hh, ww = 24, 40
    deltas = [(1, 2, 3), (-2, 0, 1)]
    a = np.zeros((hh, ww, 3), dtype=np.int64)
    mask = [[None] * ww for _ in range(hh)]
    starts = [tuple(int(v) for v in rng.integers(60, 120, size=3)) for _ in range(5)]
    for y in range(hh):
        a[y, 0] = starts[y % 5]
        for x in range(1, ww):
            if (x * 7 + y * 3) % 11 == 0:
                a[y, x] = starts[(x + y) % 5]
            else:
                d = (x + y) % 2
                a[y, x] = a[y, x - 1] + np.array(deltas[d])
                mask[y][x] = d
    assert a.min() >= 0 and a.max() <= 255
    pald = dict(begin=0, num_c=3, colors=starts, deltas=deltas, predictor=1, delta_mask=mask)
    out.append(("rgb8_palette_deltas", W.modular_image(_planes(a), tree=tree, data_code=code, palette=pald), a.astype(np.uint8), dict(width=ww, height=hh, format="Rgb", num_channels=3)))
    # 13b. A channel palette on alpha alone (begin_c = 3, num_c = 1: 5 alpha levels) after a YCgCo RCT of the colour channels; then two palettes in one
    #      image: R,G,B -> index, and alpha -> index (the second one addresses the channel list AFTER the first: palette, index, alpha -> begin 2)
    levels = [0, 60, 128, 200, 255]
    a = np.concatenate([_img(40, 52, 3, seed=18), np.array(levels)[np.random.default_rng(19).integers(0, 5, size=(40, 52, 1))]], axis=2)
    tree_a = W.Split(0, 0, W.Leaf(0, 5), W.Leaf(1, 0))
    code_a = W.EntropyCode([0, 1], [("flat", 256), ("flat", 256)], hybrids=[W.Hybrid(4, 2, 0)] * 2, log_alpha=8)
    out.append(("rgba8_rct_then_alpha_palette", W.modular_image(_planes(a), alpha_bits=8, tree=tree_a, data_code=code_a, rct=(0, 6), palette=dict(begin=3, num_c=1, colors=[(v,) for v in levels])),
                a.astype(np.uint8), dict(width=52, height=40, format="Rgb", num_channels=4, has_transparency=True)))
    a = np.concatenate([paletted(36, 44, 4), np.array(levels)[np.random.default_rng(20).integers(0, 5, size=(36, 44, 1))]], axis=2)
    out.append(("rgba8_two_palettes", W.modular_image(_planes(a), alpha_bits=8, tree=tree_a, data_code=code_a, palette=[dict(begin=0, num_c=3, colors=explicit), dict(begin=2, num_c=1, colors=[(v,) for v in levels])]),
                a.astype(np.uint8), dict(width=44, height=36, format="Rgb", num_channels=4, has_transparency=True)))
    # 13c. Local MA trees (use_global_tree = 0): (a) a frame without a global tree whose single stream brings its own; (b) a 2 x 2-group frame whose
    #      global tree is never used by data: every group section brings a different tree (predictor W + gradient split on |N|) and its own code;
    #      (c) no global tree at all, the palette in the global stream and the index groups all local
    a = _img(50, 70, 3, seed=21)
    ltree = W.Split(4, 20, W.Leaf(0, 5), W.Leaf(1, 1))                                                      # |N| > 20 ? gradient : W
    lcode = W.EntropyCode([0, 1], [("flat", 256), ("flat", 256)], hybrids=[W.Hybrid(4, 2, 0)] * 2, log_alpha=8)
    out.append(("rgb8_local_tree_single_stream", W.modular_image(_planes(a), tree=ltree, data_code=lcode, local_global=True), a.astype(np.uint8), dict(width=70, height=50, format="Rgb", num_channels=3)))
    a = _img(200, 150, 3, seed=22)
    out.append(("rgb8_local_trees_in_groups", W.modular_image(_planes(a), group_size_shift=0, group_local=(ltree, lcode)), a.astype(np.uint8), dict(width=150, height=200, format="Rgb", num_channels=3)))
    a = paletted(140, 150, 5)
    ptree = W.Split(0, 0, W.Leaf(0, 1), W.Leaf(1, 0))
    pcode = W.EntropyCode([0, 1], [("flat", 128), ("flat", 256)], hybrids=[W.Hybrid(4, 2, 0)] * 2, log_alpha=8)
    itree = W.Leaf(0, 1)
    icode = W.EntropyCode([0], [("flat", 128)], hybrids=[W.Hybrid(4, 2, 0)], log_alpha=8)
    out.append(("rgb8_palette_all_local", W.modular_image(_planes(a), tree=ptree, data_code=pcode, palette=pal, group_size_shift=0, local_global=True, group_local=(itree, icode)),
                a.astype(np.uint8), dict(width=150, height=140, format="Rgb", num_channels=3)))
    # 13d. RCTs chosen per group (listed in each group section's own header): 3 x 2 groups, a different RCT in every group, two RCTs in one of them,
    #      none in another; then the same on R,G,B of an RGBA image together with a local tree per group
    a = _img(200, 300, 3, seed=23)
    per_group = {0: [(0, 6)], 1: [(0, 10)], 2: [], 3: [(0, 17), (0, 2)], 4: [(0, 41)], 5: [(0, 1)]}
    out.append(("rgb8_rct_per_group", W.modular_image(_planes(a), group_size_shift=0, group_rct=lambda gi: per_group[gi]), a.astype(np.uint8), dict(width=300, height=200, format="Rgb", num_channels=3)))
    a = _img(150, 200, 4, seed=24)
    out.append(("rgba8_rct_per_group_local_trees", W.modular_image(_planes(a), alpha_bits=8, group_size_shift=0, group_rct=lambda gi: [(0, 6 + gi)], group_local=(ltree, lcode)),
                a.astype(np.uint8), dict(width=200, height=150, format="Rgb", num_channels=4, has_transparency=True)))
    # 13e. Palettes listed in group sections' own headers: a channel palette on alpha in every group (its own level set per group), and in one group an
    #      RGB palette of that group's colours followed by nothing, in another an RCT followed by a channel palette on the first channel
    a = _img(150, 200, 4, seed=25)
    a[..., 3] = (a[..., 3] // 40) * 40 + 7
    def gt_alpha(gi):
        x0g, y0g = (gi % 2) * 128, (gi // 2) * 128
        sub = a[y0g:y0g + 128, x0g:x0g + 128]
        lv = sorted(set(int(v) for v in sub[..., 3].ravel()))
        gts = [("palette", dict(begin=3, num_c=1, colors=[(v,) for v in lv]))]
        if gi == 1:
            gts = [(0, 6)] + gts
        return gts
    out.append(("rgba8_group_alpha_palettes", W.modular_image(_planes(a), alpha_bits=8, group_size_shift=0, group_rct=gt_alpha), a.astype(np.uint8), dict(width=200, height=150, format="Rgb", num_channels=4, has_transparency=True)))
    a = paletted(130, 250, 6)
    def gt_rgb(gi):
        x0g, y0g = (gi % 2) * 128, (gi // 2) * 128
        sub = a[y0g:y0g + 128, x0g:x0g + 128].reshape(-1, 3)
        cols = sorted(set(tuple(int(v) for v in px) for px in sub))
        return [("palette", dict(begin=0, num_c=3, colors=cols))] if gi != 2 else []
    out.append(("rgb8_group_rgb_palettes", W.modular_image(_planes(a), group_size_shift=0, group_rct=gt_rgb), a.astype(np.uint8), dict(width=250, height=130, format="Rgb", num_channels=3)))
    # 13f. Squeeze (H.6.2): the default parameter list (chroma first, then alternating steps down to 8 x 8) in a one-group frame (every channel in the
    #      global stream) and in a 2 x 2-group frame (residual channels spread over global stream and pass groups); explicit steps incl. a
    #      not-in-place one on an RGBA image; and a 1100-px-wide frame at group size 128, whose shift >= 3 channels are wider than a group and
    #      therefore live in the LF-group sections (2 LF groups of 1024 px)
    a = _img(40, 52, 3, seed=31)
    out.append(("rgb8_squeeze_default_single", W.modular_squeeze_image(_planes(a)), a.astype(np.uint8), dict(width=52, height=40, format="Rgb", num_channels=3)))
    a = _img(150, 200, 3, seed=32)
    out.append(("rgb8_squeeze_default_groups", W.modular_squeeze_image(_planes(a), group_size_shift=0), a.astype(np.uint8), dict(width=200, height=150, format="Rgb", num_channels=3)))
    a = _img(90, 70, 4, seed=33)
    out.append(("rgba8_squeeze_explicit", W.modular_squeeze_image(_planes(a), alpha_bits=8, params=[(True, True, 0, 4), (False, True, 0, 4), (True, False, 1, 2)]),
                a.astype(np.uint8), dict(width=70, height=90, format="Rgb", num_channels=4, has_transparency=True)))
    a = _img(60, 1100, 3, seed=34)
    out.append(("rgb8_squeeze_lf_sections", W.modular_squeeze_image(_planes(a), group_size_shift=0), a.astype(np.uint8), dict(width=1100, height=60, format="Rgb", num_channels=3)))
    a = _img(64, 80, 3, seed=35)
    a[..., 1] = (a[..., 0] // 2 + a[..., 1] // 2) % 256                                                     # correlated channels: the reference properties carry information
    sq_tree = W.Split(17, 0, W.Leaf(0, 5), W.Split(18, 3, W.Leaf(1, 1), W.Leaf(2, 4)))                        # previous same-size channel: value > 0 ? gradient : (|v - g| > 3 ? W : select)
    sq_code = W.EntropyCode([0, 1, 2], [("flat", 256)] * 3, hybrids=[W.Hybrid(4, 2, 0)] * 3, log_alpha=8)
    out.append(("rgb8_squeeze_prev_channel_tree", W.modular_squeeze_image(_planes(a), tree=sq_tree, data_code=sq_code), a.astype(np.uint8), dict(width=80, height=64, format="Rgb", num_channels=3)))
    a = _img(60, 1100, 3, seed=36)
    out.append(("rgb8_squeeze_lf_sections_local_trees", W.modular_squeeze_image(_planes(a), group_size_shift=0, section_local=(ltree, lcode)), a.astype(np.uint8), dict(width=1100, height=60, format="Rgb", num_channels=3)))
    # 14. Layered stills (F.2 crop + BlendingInfo, reference slots): every frame carries its crop rectangle; expected pixels from the numpy
    #     statement of the blend modes (tests/layer_util.py). (a) alpha-blend of a layer that hangs over the canvas border, (b) add, (c) three
    #     layers: the second goes to slot 2 from the empty slot 2, the last one multiplies onto slot 1
    import layer_util as LU
    cw, chh = 60, 50
    base = _img(chh, cw, 4, seed=15)
    base[..., 3] = np.maximum(base[..., 3], 90)
    over = _img(22, 28, 4, seed=16, smooth=False)
    f = lambda a: a.astype(np.float32) / 255.0
    for nm, mode, mname, (x0, y0) in (("layers_blend_over_border", 2, "blend", (45, -6)), ("layers_add", 1, "add", (10, 8))):
        data = W.modular_layers(cw, chh, [dict(channels=_planes(base)), dict(channels=_planes(over), x0=x0, y0=y0, blend=dict(mode=mode, alpha_channel=0, source=0))])
        want = LU.to_u8(LU.composite(f(base), f(over), x0, y0, mname))
        out.append((nm, data, ("within1", want), dict(width=cw, height=chh, format="Rgb", num_channels=4, has_transparency=True)))
    other = _img(chh, cw, 4, seed=17)
    data = W.modular_layers(cw, chh, [dict(channels=_planes(base), save=1), dict(channels=_planes(other), blend=dict(mode=1, source=2), save=2),
                                      dict(channels=_planes(over), x0=5, y0=20, blend=dict(mode=4, source=1, clamp=True))])
    want = LU.to_u8(LU.composite(f(base), f(over), 5, 20, "mul"))
    out.append(("layers_three_slots_mul", data, ("within1", want), dict(width=cw, height=chh, format="Rgb", num_channels=4, has_transparency=True)))
    return out



def containerised():
    """Container-level variants of case 2: jxlp split with metadata boxes between the parts, a jxll level box."""
    name, cs, px, info = cases()[1]
    exif = b"\x00\x00\x00\x00II*\x00\x08\x00\x00\x00\x00\x00"
    xmp = b"<x:xmpmeta xmlns:x='adobe:ns:meta/'/>"
    return [
        ("jxlc_with_boxes", W.container(cs, boxes=[(b"Exif", exif), (b"xml ", xmp)], level=5), px, exif, [xmp]),
        ("jxlp_split", W.container(cs, boxes=[(b"Exif", exif), (b"xml ", xmp), (b"xml ", xmp + b"2")], split_at=len(cs) // 3), px, exif, [xmp, xmp + b"2"]),
    ]


# ---- DC-only VarDCT frame: expected pixels from the published constants, computed here in float64
OPSIN_INV = np.array([[11.031566901960783, -9.866943921568629, -0.16462299647058826],
                      [-3.254147380392157, 4.418770392156863, -0.16462299647058826],
                      [-3.6588512862745097, 2.7129230470588235, 1.9459282392156863]])
OPSIN_BIAS = 0.0037930732552754493


def vardct_dc_case(seed=11, xb=9, yb=7, xsize=70, ysize=52):
    rng = np.random.default_rng(seed)
    gs, qlf = 32768, 64
    yy, xx = np.mgrid[0:yb, 0:xb]
    qy = (40 + 6 * xx + 3 * yy + rng.integers(-3, 4, size=(yb, xb))).astype(np.int64)      # Y in about [0.2, 0.7]
    qx = rng.integers(-6, 7, size=(yb, xb)).astype(np.int64)
    qb = rng.integers(-20, 21, size=(yb, xb)).astype(np.int64)
    f = W.vardct_dc_only([_planes(qx)[0], _planes(qy)[0], _planes(qb)[0]], xsize, ysize, global_scale=gs, quant_lf=qlf)
    inv = 65536.0 / gs / qlf
    Y = qy * (inv / 512)
    X = qx * (inv / 4096)             # default chroma-from-luma: X gets 0 * Y, B gets 1 * Y
    B = qb * (inv / 256) + Y
    cb = np.cbrt(OPSIN_BIAS)
    mix = np.stack([(Y + X + cb) ** 3 - OPSIN_BIAS, (Y - X + cb) ** 3 - OPSIN_BIAS, (B + cb) ** 3 - OPSIN_BIAS], axis=-1)
    lin = mix @ OPSIN_INV.T
    a = np.abs(lin)
    srgb = np.where(a <= 0.0031308, 12.92 * a, 1.055 * np.power(a, 1 / 2.4) - 0.055) * np.sign(lin)
    block = np.clip(srgb, 0, 1) * 255.0
    full = np.repeat(np.repeat(block, 8, axis=0), 8, axis=1)[:ysize, :xsize]
    return f, full      # float expected values: compare with |decoded - expected| <= 0.5 + eps
