"""numpy statement of the frame-header blend modes (replace / add / blend / alpha-weighted add / multiply) on float RGBA canvases:
the reference the layer tests compare the oracle and the GPU decoder with."""
import numpy as np


def composite(bg, fg, x0, y0, mode, premultiplied=False):
    """float RGBA canvas `bg` (H x W x 4, 0..1) with layer `fg` blended in at (x0, y0); the formulas of the frame header's blend modes."""
    out = bg.copy()
    ys, xs = np.mgrid[0:fg.shape[0], 0:fg.shape[1]]
    cy, cx = ys + y0, xs + x0
    ok = (cy >= 0) & (cy < bg.shape[0]) & (cx >= 0) & (cx < bg.shape[1])
    f, b = fg[ys[ok], xs[ok]], bg[cy[ok], cx[ok]]
    fa, ba = f[:, 3:4], b[:, 3:4]
    if mode == "replace":
        o = f
    elif mode == "add":
        o = b + f
    elif mode == "blend":
        na = 1 - (1 - fa) * (1 - ba)
        if premultiplied:
            col = f[:, :3] + b[:, :3] * (1 - fa)
        else:
            col = (f[:, :3] * fa + b[:, :3] * ba * (1 - fa)) * np.where(na > 0, 1 / np.maximum(na, 1e-30), 0)
        o = np.concatenate([col, na], axis=1)
    elif mode == "muladd":
        o = np.concatenate([b[:, :3] + f[:, :3] * fa, ba], axis=1)
    else:
        o = np.concatenate([b[:, :3] * f[:, :3], ba * fa], axis=1)
    out[cy[ok], cx[ok]] = o
    return out


def to_u8(canvas):
    return np.clip(np.rint(canvas * 255.0), 0, 255).astype(np.uint8)
