"""Seeded inputs shared by the CPU and GPU tests."""
import numpy as np


def gradient(w, h, c, seed, sigma=4.0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    ramp = 255.0 * (x + y) / max(w + h - 2, 1)
    img = np.clip(ramp[..., None] + 20.0 * np.arange(c) + rng.normal(0, sigma, (h, w, c)), 0, 255).astype(np.uint8)
    if c == 4:
        img[..., 3] = 255
    return img


def noise(w, h, c, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, c), dtype=np.uint8)


def flat(w, h, c, v=7):
    return np.full((h, w, c), v, dtype=np.uint8)


def skewed(w, h, c, seed):
    """Geometric residuals: drives Huffman depths past the 11-bit cap so the Kraft repair runs."""
    rng = np.random.default_rng(seed)
    steps = (rng.geometric(0.55, size=(h, w, c)) - 1).astype(np.int64)
    steps *= rng.choice([-1, 1], size=steps.shape)
    return (np.cumsum(steps, axis=1) & 255).astype(np.uint8)


def with_const(img, **chans):
    """Force channels to constants (flat channels, FLP0 §2b); e.g. with_const(img, c1=77)."""
    img = img.copy()
    for k, v in chans.items():
        img[..., int(k[1:])] = v
    return img


def alpha_ramp(w, h, seed):
    img = gradient(w, h, 4, seed)
    img[..., 3] = (np.arange(w)[None, :] * 255 // max(w - 1, 1)).astype(np.uint8)
    return img


# flags: predictor 1 | 0x10 subtract-green | 0x20 one stream per block | 0x40 exact block sizes
ALL_FLAGS = [0x01, 0x11, 0x21, 0x31, 0x41, 0x51]

# (name, builder) — sizes chosen so the CPU model finishes each in well under a second
SMALL = [
    ("c1_512x512x3", lambda: gradient(512, 512, 3, 1)),
    ("rgba_256x64", lambda: gradient(256, 64, 4, 2)),
    ("ragged_130x33x4", lambda: gradient(130, 33, 4, 3)),
    ("ragged_127x31x3", lambda: gradient(127, 31, 3, 4)),
    ("odd_pitch_131x70x3", lambda: gradient(131, 70, 3, 5)),
    ("gray_200x40x1", lambda: gradient(200, 40, 1, 6)),
    ("ga_77x45x2", lambda: gradient(77, 45, 2, 7)),
    ("one_pixel", lambda: gradient(1, 1, 1, 8)),
    ("one_row_300x1x4", lambda: gradient(300, 1, 4, 9)),
    ("one_col_1x100x3", lambda: gradient(1, 100, 3, 10)),
    ("noise_256x96x4", lambda: noise(256, 96, 4, 11)),
    ("noise_129x65x3", lambda: noise(129, 65, 3, 12)),
    ("flat_256x64x4", lambda: flat(256, 64, 4)),
    ("flat_100x50x3", lambda: flat(100, 50, 3, 200)),
    ("skewed_384x96x4", lambda: skewed(384, 96, 4, 13)),
    ("skewed_128x32x1", lambda: skewed(128, 32, 1, 14)),
    # flat channels: none (alpha ramp), green only, alpha of gray+alpha, two of four, flat in some blocks only
    ("alpha_ramp_256x64x4", lambda: alpha_ramp(256, 64, 15)),
    ("const_g_200x40x3", lambda: with_const(gradient(200, 40, 3, 16), c1=77)),
    ("const_a_77x45x2", lambda: with_const(gradient(77, 45, 2, 17), c1=255)),
    ("two_flat_300x70x4", lambda: with_const(gradient(300, 70, 4, 18), c2=9, c3=200)),
    ("partly_flat_384x64x4", lambda: np.concatenate([gradient(128, 64, 4, 19), noise(256, 64, 4, 20)], axis=1)),
]
