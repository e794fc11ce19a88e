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


def from_residuals(res):
    """The image whose FLP0 residuals (left predictor; column 0 from the pixel above; no colour transform) are `res`."""
    res = res.astype(np.int64)
    col0 = np.cumsum(res[:, 0, :], axis=0)                      # column 0: running sum down the rows
    rows = np.cumsum(res[:, 1:, :], axis=1) + col0[:, None, :]  # then along each row
    return (np.concatenate([col0[:, None, :], rows], axis=1) & 255).astype(np.uint8)


def max_len_row(seed, c=4):
    """One block whose LAST row is made of maximum-length codes only (128*c symbols x 10 bits = the 160-word worst case
    of a row sub-stream when c == 4): six frequent symbols in rows 0..30, 250 rare ones in row 31."""
    rng = np.random.default_rng(seed)
    n = 31 * 128 * c
    counts = np.array([0.504, 0.252, 0.126, 0.063, 0.0315]) * n
    counts = counts.astype(np.int64)
    top = np.repeat(np.arange(6), list(counts) + [n - counts.sum()])
    rng.shuffle(top)
    rare = np.concatenate([np.arange(6, 256), rng.integers(6, 256, 128 * c)])[:128 * c] if 128 * c >= 250 else np.arange(6, 6 + 128 * c)
    rng.shuffle(rare)
    res = np.concatenate([top, rare]).reshape(32, 128, c)
    return from_residuals(res)


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
    # the last row of the block is 160 words long (every symbol a 10-bit code): the staging tile's very last word
    ("max_len_row_128x32x4", lambda: max_len_row(21, 4)),
    ("max_len_row_128x32x3", lambda: max_len_row(22, 3)),
]
