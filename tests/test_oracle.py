"""CPU tests of the FLP0 model (oracle/).  These pin the provisional format against committed
self-goldens; they do NOT pin anything against the reference (licensing gate — LICENSING.md)."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,build", cases.SMALL, ids=[n for n, _ in cases.SMALL])
@pytest.mark.parametrize("flags", cases.ALL_FLAGS)
def test_roundtrip(oracle, name, build, flags):
    img = build()
    s = oracle.encode(img, flags)
    assert np.array_equal(oracle.decode(s, img.shape), img)


def test_golden_streams(oracle):
    """Committed FLP0 vectors (tests/golden/make_golden.py made them from this same model)."""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        gold = json.load(f)
    builders = dict(cases.SMALL)
    for key, want in gold["sha256"].items():
        name, flags = key.rsplit("@", 1)
        s = oracle.encode(builders[name](), int(flags, 16))
        assert len(s) == want["bytes"], key
        assert hashlib.sha256(s.tobytes()).hexdigest() == want["sha256"], key
    img = np.load(os.path.join(GOLDEN, "tiny_96x40x3.npy"))
    stream = np.fromfile(os.path.join(GOLDEN, "tiny_96x40x3.flp0"), dtype=np.uint8)
    assert np.array_equal(oracle.encode(img, 0x01), stream)
    assert np.array_equal(oracle.decode(stream, img.shape), img)


def test_flat_channels(oracle):
    """FLP0 §2b: a channel that is constant over a block is named in the block header and costs no bits."""
    img = cases.gradient(256, 64, 4, 2)                      # alpha = 255 everywhere
    _, meta = oracle.block_histograms(img, with_flat=True)
    assert (meta[:, 0] == 8).all() and (meta[:, 1] == 0xFF000000).all()
    ramp = cases.alpha_ramp(256, 64, 2)
    _, meta = oracle.block_histograms(ramp, with_flat=True)
    assert (meta[:, 0] == 0).all()
    assert len(oracle.encode(img)) < len(oracle.encode(ramp))
    flat = cases.flat(256, 64, 4)                            # every channel flat: headers only
    h, meta = oracle.block_histograms(flat, with_flat=True)
    assert (meta[:, 0] == 15).all() and h.sum() == 0
    assert len(oracle.encode(flat)) == 32 + 4 * (4 + 1) + 4 * 4 * 50
    # with subtract-green it is the TRANSFORMED value that must be constant
    sg = cases.with_const(cases.gradient(128, 32, 3, 3), c0=10)
    assert oracle.block_histograms(sg, 0x01, with_flat=True)[1][0, 0] == 1
    assert oracle.block_histograms(sg, 0x11, with_flat=True)[1][0, 0] == 0


def test_known_answer_lengths(oracle):
    """Hand-checkable Huffman cases for the length builder."""
    h = np.zeros(256, dtype=np.uint32)
    assert oracle.lengths(h).sum() == 0                      # empty
    h[42] = 9
    ln = oracle.lengths(h)
    assert ln[42] == 15 and ln.sum() == 15                   # sole symbol: zero-length marker
    h[:] = 0
    h[[1, 2]] = [5, 3]
    assert list(oracle.lengths(h)[[1, 2]]) == [1, 1]
    h[:] = 0
    h[[10, 20, 30, 40]] = [8, 4, 2, 1]                       # 1,2,3,3
    assert list(oracle.lengths(h)[[10, 20, 30, 40]]) == [1, 2, 3, 3]
    h[:] = 1                                                 # uniform 256 -> all 8
    assert set(oracle.lengths(h)) == {8}


def test_length_limit_and_kraft(oracle):
    """Fibonacci counts force depth > 10; the repair must keep Kraft equality and the cap."""
    fib = [1, 1]
    while len(fib) < 20:
        fib.append(fib[-1] + fib[-2])
    h = np.zeros(256, dtype=np.uint32)
    h[:20] = fib
    ln = oracle.lengths(h).astype(int)
    used = ln[ln > 0]
    assert used.max() == 10
    assert sum(2 ** (10 - l) for l in used) == 2 ** 10
    # rarer symbols never get shorter codes than more frequent ones
    order = np.argsort(h[:20], kind="stable")
    assert all(ln[order[i]] >= ln[order[i + 1]] for i in range(19))
    rng = np.random.default_rng(0)
    for _ in range(200):
        h = (rng.geometric(rng.uniform(0.01, 0.9), 256) * (rng.random(256) < rng.uniform(0.05, 1))).astype(np.uint32)
        ln = oracle.lengths(h).astype(int)
        used = ln[(ln > 0) & (ln != 15)]
        if used.size:
            assert used.max() <= 11 and sum(2 ** (10 - l) for l in used) == 2 ** 10


def test_canonical_codes_prefix_free(oracle):
    rng = np.random.default_rng(1)
    h = rng.integers(0, 500, 256).astype(np.uint32)
    t = oracle.table(h).astype(int)
    codes = sorted((format(e & 0xFFF, "b").zfill(e >> 12) for e in t if 1 <= (e >> 12) <= 10))
    assert all(not b.startswith(a) for a, b in zip(codes, codes[1:]))


def test_errors(oracle):
    img = cases.gradient(64, 64, 3, 1)
    assert oracle.encode_rc(img, flags=0x02) == -1           # unknown predictor
    assert oracle.encode_rc(img, flags=0x81) == -1           # reserved flag bit
    assert oracle.encode_rc(img, flags=0x61) == -1           # ONE_STREAM and EXACT cannot be combined
    assert oracle.encode_rc(img, cap=100) == -2              # capacity
    s = oracle.encode(img)
    assert oracle.decode_rc(s[:20], img.shape) == -3         # truncated header
    assert oracle.decode_rc(s[:-8], img.shape) == -3         # truncated payload
    bad = s.copy(); bad[0] ^= 0xFF
    assert oracle.decode_rc(bad, img.shape) == -3            # magic
    bad = s.copy(); bad[32 + 4] = 0xFF; bad[32 + 7] = 0x7F   # directory entry beyond payload
    assert oracle.decode_rc(bad, img.shape) == -3
    assert oracle.decode_rc(s, (64, 64, 2)) == -2            # output too small


def test_blocks_are_independent(oracle):
    """A block-row slice encodes to the same block payloads as inside the full image — the
    property the multi-GPU block-row split relies on."""
    img = cases.gradient(300, 100, 3, 21)
    full = oracle.encode(img)
    top, bot = oracle.encode(img[:64]), oracle.encode(img[64:])
    nb_t = int(np.frombuffer(top[20:24], np.uint32)[0]); nb_b = int(np.frombuffer(bot[20:24], np.uint32)[0])
    pay = lambda s, nb: s[32 + 4 * (nb + 1):]
    assert np.array_equal(np.concatenate([pay(top, nb_t), pay(bot, nb_b)]), pay(full, nb_t + nb_b))


def test_layout_modes(oracle):
    """FLP0 §8: EXACT drops the slot slack and nothing else; ONE_STREAM drops the row word counts, the row
    padding and the slack.  All three layouts carry the same symbols with the same code tables."""
    img = cases.gradient(300, 100, 3, 31)
    slot, exact, one = oracle.encode(img, 0x01), oracle.encode(img, 0x41), oracle.encode(img, 0x21)
    assert one.size < exact.size < slot.size
    nb = int(np.frombuffer(slot[20:24], np.uint32)[0])
    d = lambda s: np.frombuffer(s[32:32 + 4 * (nb + 1)].tobytes(), np.uint32).astype(np.int64)
    ds, de, do = d(slot), d(exact), d(one)
    pay = lambda s: np.frombuffer(s[32 + 4 * (nb + 1):].tobytes(), np.uint32)
    for b in range(nb):
        bs, be, bo = pay(slot)[ds[b]:ds[b + 1]], pay(exact)[de[b]:de[b + 1]], pay(one)[do[b]:do[b + 1]]
        assert np.array_equal(bs[:32], be[:32]) and np.array_equal(bs[:32], bo[:32])      # length nibbles
        assert np.array_equal(bs[:be.size], be) and not bs[be.size:].any()                # slot = exact + zero slack
        rows = np.frombuffer(be[32:48].tobytes(), np.uint16).astype(np.int64)
        assert be.size == 50 + rows.sum()
        assert np.array_equal(be[48:50], bo[32:34])                                       # flat words
        assert 0 <= rows.sum() - (bo.size - 34) <= 32                                     # at most one pad word per row saved
    assert int(slot[7]) == 0x01 and int(exact[7]) == 0x41 and int(one[7]) == 0x21        # the header says which
