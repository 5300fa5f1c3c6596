"""Deterministic synthetic test images (SURVEY.md §8d recipe): bilinear colour gradients + 1/f noise + hard-edged shapes.
Pure numpy; shared by tests/, bench.py and scripts/ (it is not part of the oracle and not part of the engine)."""
import numpy as np


def synthetic_image(w, h, seed=0, channels=3):
    """Deterministic photo-like test image (SURVEY.md §8d recipe): gradients + 1/f noise + hard-edged shapes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.meshgrid(np.linspace(0, 1, h, dtype=np.float32), np.linspace(0, 1, w, dtype=np.float32), indexing="ij")
    corners = rng.random((4, 3)).astype(np.float32)
    L = ((1 - yy)[..., None] * ((1 - xx)[..., None] * corners[0] + xx[..., None] * corners[1]) +
         yy[..., None] * ((1 - xx)[..., None] * corners[2] + xx[..., None] * corners[3]))
    fy = np.fft.fftfreq(h)[:, None] * h
    fx = np.fft.rfftfreq(w)[None, :] * w
    f = np.sqrt(fy * fy + fx * fx)
    filt = 1.0 / np.maximum(f, 1.0 / 64)
    base = rng.standard_normal((h, w)).astype(np.float32)
    N = np.empty((h, w, 3), np.float32)
    for c in range(3):
        white = 0.6 * base + 0.4 * rng.standard_normal((h, w)).astype(np.float32)
        n = np.fft.irfft2(np.fft.rfft2(white) * filt, s=(h, w)).astype(np.float32)
        n -= n.min()
        n /= max(float(n.max()), 1e-9)
        N[..., c] = n
    E = np.zeros((h, w, 3), np.float32)
    for _ in range(24):
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        x1, y1 = min(w, x0 + int(rng.integers(4, max(5, w // 4)))), min(h, y0 + int(rng.integers(4, max(5, h // 4))))
        E[y0:y1, x0:x1] = rng.random(3)
    Y, X = np.ogrid[:h, :w]
    for _ in range(12):
        cx, cy, r = int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(3, max(4, min(w, h) // 8)))
        E[(X - cx) ** 2 + (Y - cy) ** 2 <= r * r] = rng.random(3)
    img = np.clip(0.55 * L + 0.30 * N + 0.15 * E, 0, 1)
    out = np.round(255 * img).astype(np.uint8)
    if channels == 1:
        return out[..., 1:2].copy()
    if channels == 4:
        r2 = ((xx - 0.5) ** 2 + (yy - 0.5) ** 2) * 4
        a = np.clip(1.6 - 1.8 * r2 + 0.3 * (rng.random((h, w)).astype(np.float32) - 0.5), 0, 1)
        a[a > 0.62] = 1.0
        a[a < 0.05] = 0.0
        return np.concatenate([out, np.round(255 * a).astype(np.uint8)[..., None]], axis=2)
    return out


