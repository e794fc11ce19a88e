"""Synthetic workloads C1..C5 (BASELINE.json `configs`, SURVEY.md §8d / BASELINE.md).

Generator: numpy PCG64 (`default_rng(seed)`), fixed seeds.  One deliberate
deviation from the survey's recipe, stated here and in DESIGN.md: the survey's
ramp `(x+y)/4` saturates at 255 beyond x+y = 1020, which would make most of a
4K image a flat white field (trivially compressible).  The ramp here spans the
image diagonal (0..255 corner to corner) so every pixel carries gradient plus
N(0, sigma=4) noise — a harder, more honest input.
"""
import numpy as np


def gradient_noise(w, h, c, seed, sigma=4.0, alpha="opaque"):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    ramp = 255.0 * (x + y) / max(w + h - 2, 1)
    img = np.empty((h, w, c), dtype=np.uint8)
    for ch in range(c):
        if ch == 3:
            img[..., 3] = 255 if alpha == "opaque" else np.clip(255.0 - ramp, 0, 255).astype(np.uint8)
            continue
        base = ramp if ch != 1 else 255.0 - ramp  # green runs the other way
        noise = rng.normal(0.0, sigma, size=(h, w)).astype(np.float32)
        img[..., ch] = np.clip(base + noise, 0, 255).astype(np.uint8)
    return img


def gradient_noise_rows(w, h, c, seed, y0, y1, sigma=4.0, band=32):
    """Rows [y0, y1) of a w x h gradient+noise image whose noise is seeded per 32-row band, so that
    any rank can synthesise exactly its own block rows of one huge image (config C4) without
    generating the rest.  Not the same pixels as gradient_noise(); same statistics."""
    out = np.empty((y1 - y0, w, c), dtype=np.uint8)
    x = np.arange(w, dtype=np.float32)[None, :]
    for b0 in range(y0 - y0 % band, y1, band):
        lo, hi = max(b0, y0), min(b0 + band, y1, h)
        if hi <= lo:
            continue
        rng = np.random.default_rng([seed, b0 // band])
        noise = rng.normal(0.0, sigma, size=(band, w, min(c, 3))).astype(np.float32)
        y = np.arange(lo, hi, dtype=np.float32)[:, None]
        ramp = 255.0 * (x + y) / max(w + h - 2, 1)
        for ch in range(c):
            if ch == 3:
                out[lo - y0:hi - y0, :, 3] = 255
                continue
            base = ramp if ch != 1 else 255.0 - ramp
            out[lo - y0:hi - y0, :, ch] = np.clip(base + noise[lo - b0:hi - b0, :, ch], 0, 255).astype(np.uint8)
    return out


def uniform_noise(w, h, c, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, c), dtype=np.uint8)


CONFIGS = {
    # id: (n_images, w, h, c, kind, first_seed)
    "C1": (1, 512, 512, 3, "gradient", 1),
    "C2": (1, 3840, 2160, 4, "gradient", 2),
    "C2A": (1, 3840, 2160, 4, "gradient-alpha", 2),  # SURVEY §8(d)'s second variant: alpha is a gradient too
    "C3": (1024, 1920, 1080, 3, "gradient", 1000),
    "C4": (1, 16384, 16384, 4, "gradient", 4),
    "C5": (64, 3840, 2160, 4, "uniform", 5000),
}


def make_batch(cfg, n=None, distinct=8):
    """Batch for config `cfg` as uint8 [n,h,w,c].  Only `distinct` different images are
    synthesised (seeds first_seed..); the rest of the batch cycles through them, which
    keeps host-side generation to seconds without changing per-image statistics."""
    n_cfg, w, h, c, kind, seed0 = CONFIGS[cfg]
    n = n_cfg if n is None else n
    k = min(n, distinct)
    if kind == "gradient":
        gen = lambda s: gradient_noise(w, h, c, s)
    elif kind == "gradient-alpha":
        gen = lambda s: gradient_noise(w, h, c, s, alpha="ramp")
    else:
        gen = lambda s: uniform_noise(w, h, c, s)
    uniq = [gen(seed0 + i) for i in range(k)]
    out = np.empty((n, h, w, c), dtype=np.uint8)
    for i in range(n):
        out[i] = uniq[i % k]
    return out
