"""GPU parity tests (run with -m gpu on a B200): the CUDA engine, called through the C ABI,
against the CPU model of the provisional FLP0 format — byte-exact streams, pixel-exact decodes,
stage-by-stage (histograms, code tables), plus size-independent round-trip properties at
BASELINE.json's full sizes.  "Parity" here is engine-vs-model; parity with the reference is
unpinned (licensing gate, LICENSING.md)."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name,build", cases.SMALL, ids=[n for n, _ in cases.SMALL])
@pytest.mark.parametrize("flags", [0x01, 0x11])
def test_stage_histograms_and_tables(codec, oracle, name, build, flags):
    img = build()
    px = dev(img[None])
    nb = -(-img.shape[1] // 128) * -(-img.shape[0] // 32)
    hist = torch.zeros((nb, 256), dtype=torch.int16, device="cuda")
    table = torch.zeros((nb, 256), dtype=torch.int16, device="cuda")
    flat = torch.zeros((nb, 2), dtype=torch.int32, device="cuda")
    codec.stage_histograms(px, hist, flags, flat=flat)
    bits = torch.zeros(nb, dtype=torch.int32, device="cuda")
    codec.stage_tables(hist, table, bits=bits)
    codec.check()
    want_h, want_f = oracle.block_histograms(img, flags, with_flat=True)
    got_h = hist.cpu().numpy().view(np.uint16)
    assert np.array_equal(flat.cpu().numpy().view(np.uint32), want_f), "flat-channel mask / values"
    assert np.array_equal(got_h, want_h), f"first bad block {np.argwhere((got_h != want_h).any(1))[:1]}"
    got_t = table.cpu().numpy().view(np.uint16)
    got_b = bits.cpu().numpy()
    for b in range(nb):
        want_t = oracle.table(want_h[b])
        assert np.array_equal(got_t[b], want_t), f"block {b}"
        ln = (want_t >> 12).astype(np.int64)
        assert got_b[b] == int((want_h[b].astype(np.int64) * np.where(ln <= 10, ln, 0)).sum()), f"code bits of block {b}"


@pytest.mark.parametrize("name,build", cases.SMALL, ids=[n for n, _ in cases.SMALL])
@pytest.mark.parametrize("flags", [0x01, 0x11])
def test_encode_bytes_and_decode_pixels(codec, oracle, name, build, flags):
    img = build()
    want = oracle.encode(img, flags)
    got = codec.encode(img, flags)
    assert got.size == want.size, (got.size, want.size)
    assert np.array_equal(got, want), f"first differing byte {int(np.argmax(got != want))}"
    assert np.array_equal(codec.decode(want), img)          # GPU decodes the model's stream
    assert np.array_equal(oracle.decode(got, img.shape), img)  # model decodes the GPU's stream


def test_golden_vectors(codec):
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        gold = json.load(f)
    builders = dict(cases.SMALL)
    for key, want in gold["sha256"].items():
        name, flags = key.rsplit("@", 1)
        s = codec.encode(builders[name](), int(flags, 16))
        assert s.size == want["bytes"] and hashlib.sha256(s.tobytes()).hexdigest() == want["sha256"], key
    img = np.load(os.path.join(GOLDEN, "tiny_96x40x3.npy"))
    stream = np.fromfile(os.path.join(GOLDEN, "tiny_96x40x3.flp0"), dtype=np.uint8)
    assert np.array_equal(codec.encode(img), stream)
    assert np.array_equal(codec.decode(stream), img)


def test_batch_matches_per_image(codec, oracle):
    """A batch is the back-to-back concatenation of the per-image streams."""
    imgs = np.stack([cases.gradient(300, 70, 3, s) for s in range(5)])
    streams, off = codec.encode_batch(imgs)
    assert off[0] == 0 and off[-1] == streams.size
    for i in range(5):
        assert np.array_equal(streams[int(off[i]): int(off[i + 1])], oracle.encode(imgs[i]))
    assert np.array_equal(codec.decode_batch(streams, off), imgs)


def test_rgba_batch_ragged_edges(codec, oracle):
    """RGBA batch whose blocks hang over the right and bottom image edges: the TMA store tiles of the
    decoder must clip there (3-D tensor map: row bytes, rows, images) and never touch the next image."""
    imgs = np.stack([cases.gradient(260, 70, 4, 200 + s) for s in range(5)])
    imgs[1, ..., 3] = (np.arange(260)[None, :] % 256).astype(np.uint8)   # one image without a flat alpha
    streams, off = codec.encode_batch(imgs)
    for i in range(5):
        assert np.array_equal(streams[int(off[i]): int(off[i + 1])], oracle.encode(imgs[i]))
    guard = np.full((7, 70, 260, 4), 0xA5, dtype=np.uint8)
    out = torch.from_numpy(guard).cuda()
    d_s, d_o = dev(streams), torch.from_numpy(off.astype(np.int64)).cuda()
    codec.decode_batch_device(d_s, d_o, out[1:6])
    codec.check()
    got = out.cpu().numpy()
    assert np.array_equal(got[1:6], imgs)
    assert (got[0] == 0xA5).all() and (got[6] == 0xA5).all()   # neighbours untouched


def test_host_pipeline_many_chunks(codec, oracle, monkeypatch):
    """The host-buffer API streams the batch through double-buffered chunks; force 1-2 images per chunk
    so buffer reuse (chunk k vs k-2) and the offset rebasing are exercised, odd tail chunk included."""
    imgs = np.stack([cases.gradient(300, 70, 3, 100 + s) for s in range(11)])
    monkeypatch.setenv("FLIC_CHUNK_BYTES", str(2 * imgs[0].nbytes))
    streams, off = codec.encode_batch(imgs)
    want = np.concatenate([oracle.encode(im) for im in imgs])
    assert off[-1] == want.size and np.array_equal(streams, want)
    assert np.array_equal(codec.decode_batch(streams, off), imgs)
    monkeypatch.setenv("FLIC_CHUNK_BYTES", "1")  # one image per chunk
    s1, o1 = codec.encode_batch(imgs)
    assert np.array_equal(s1, want) and np.array_equal(o1, off)
    assert np.array_equal(codec.decode_batch(s1, o1), imgs)


def test_unaligned_device_pointer(codec, oracle):
    """Pixels at an address that is not 16-byte aligned take the byte-granular load/store path."""
    img = cases.gradient(256, 64, 4, 41)
    raw = torch.zeros(img.size + 64, dtype=torch.uint8, device="cuda")
    px = raw[3: 3 + img.size].view(1, *img.shape)
    px.copy_(dev(img[None]))
    cap = codec.lib.flic_max_stream_bytes(256, 64, 4)
    streams = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    off = torch.zeros(2, dtype=torch.int64, device="cuda")
    codec.encode_batch_device(px, streams, off)
    codec.check()
    n = int(off[1])
    assert np.array_equal(streams[:n].cpu().numpy(), oracle.encode(img))
    out_raw = torch.zeros(img.size + 64, dtype=torch.uint8, device="cuda")
    out = out_raw[5: 5 + img.size].view(1, *img.shape)
    codec.decode_batch_device(streams, off, out)
    codec.check()
    assert np.array_equal(out.cpu().numpy()[0], img)


def test_errors_through_the_abi(codec, oracle):
    import flic_b200 as flic
    img = cases.gradient(256, 64, 3, 42)
    with pytest.raises(flic.FlicError) as e:
        codec.encode(img, flags=0x02)
    assert e.value.code == -1
    with pytest.raises(flic.FlicError) as e:
        codec.encode_batch(img[None], out=np.empty(64, np.uint8))
    assert e.value.code == -2
    s = oracle.encode(img)
    bad = s.copy(); bad[32 + 4] = 0xFF; bad[32 + 7] = 0x7F  # directory entry beyond the payload
    with pytest.raises(flic.FlicError) as e:
        codec.decode(bad)
    assert e.value.code == -3
    # a flat-channel mask naming a channel the image does not have (block header word 48 of block 0)
    nb = int(np.frombuffer(s[20:24], np.uint32)[0])
    first_block = 32 + 4 * (nb + 1) + 4 * int(np.frombuffer(s[32:36], np.uint32)[0])
    bad = s.copy(); bad[first_block + 4 * 48] = 0x08
    with pytest.raises(flic.FlicError) as e:
        codec.decode(bad)
    assert e.value.code == -3
    assert oracle.decode_rc(bad, img.shape) == -3
    # device capacity overrun is caught by the kernels, not by an out-of-bounds write
    px = dev(cases.noise(256, 64, 4, 1)[None])
    small = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    off = torch.zeros(2, dtype=torch.int64, device="cuda")
    codec.encode_batch_device(px, small, off)
    with pytest.raises(flic.FlicError) as e:
        codec.check()
    assert e.value.code == -2
    assert np.array_equal(codec.decode(codec.encode(img)), img)  # context still healthy


def test_random_geometries(codec, oracle):
    """Differential sweep: random sizes (ragged edges on both axes, widths that do and do not allow the
    aligned / TMA store paths), channel counts, colour transform, flat and partly flat channels, smooth /
    noisy / skewed content — GPU bytes == model bytes and both decoders return the pixels."""
    rng = np.random.default_rng(2024)
    for trial in range(120):
        c = int(rng.integers(1, 5))
        w = int(rng.choice([1, 7, 60, 127, 128, 129, 200, 255, 256, 260, 384, 500]))
        h = int(rng.choice([1, 5, 31, 32, 33, 64, 70]))
        kind = trial % 4
        if kind == 0:
            img = cases.gradient(w, h, c, 1000 + trial)
        elif kind == 1:
            img = cases.noise(w, h, c, 1000 + trial)
        elif kind == 2:
            img = cases.skewed(w, h, c, 1000 + trial)
        else:
            img = cases.gradient(w, h, c, 1000 + trial, sigma=float(rng.choice([0.0, 0.7, 12.0])))
        if rng.random() < 0.4:                       # make some channels flat, everywhere or in the left part only
            ch = int(rng.integers(0, c))
            x1 = w if rng.random() < 0.5 else max(1, w // 2)
            img = img.copy(); img[:, :x1, ch] = int(rng.integers(0, 256))
        flags = 0x11 if (c >= 3 and rng.random() < 0.5) else 0x01
        want = oracle.encode(img, flags)
        got = codec.encode(img, flags)
        assert np.array_equal(got, want), (trial, w, h, c, flags, kind)
        assert np.array_equal(codec.decode(got), img), (trial, w, h, c, flags, kind)
        assert np.array_equal(oracle.decode(got, img.shape), img), (trial, w, h, c, flags, kind)


def test_corrupt_streams_never_crash(codec, oracle):
    """Random damage anywhere in a stream — directory, length nibbles, row word counts, flat words, payload:
    the decoder answers OK (garbage pixels) or FLIC_E_FORMAT, never faults, and the context stays usable.
    Bounds come from validated header fields only, so no bit pattern can steer a load or store outside
    the stream / the image."""
    import flic_b200 as flic
    rng = np.random.default_rng(77)
    img = cases.gradient(384, 96, 4, 44)
    good = codec.encode(img)
    outcomes = {0: 0, -3: 0}
    for trial in range(60):
        bad = good.copy()
        lo = 0 if trial % 3 else 32            # a third of the trials spare the file header
        for _ in range(int(rng.integers(1, 12))):
            i = int(rng.integers(lo, bad.size))
            bad[i] ^= np.uint8(rng.integers(1, 256))
        try:
            out = codec.decode(bad)
            assert out.shape == img.shape
            outcomes[0] += 1
        except flic.FlicError as e:
            assert e.code == -3, e
            outcomes[-3] += 1
        rc = oracle.decode_rc(bad, img.shape)  # the CPU model must survive the same input
        assert rc in (0, -3)
    assert outcomes[0] and outcomes[-3]        # both outcomes occur, so the sweep is not vacuous
    assert np.array_equal(codec.decode(good), img)


def test_splice_of_gpu_parts(codec):
    """Block-row split: GPU-encoded halves splice into the GPU-encoded whole (the C4 multi-GPU path)."""
    import flic_b200 as flic
    img = cases.gradient(700, 200, 4, 43)
    whole = codec.encode(img)
    parts = [codec.encode(img[:96]), codec.encode(img[96:160]), codec.encode(img[160:])]
    assert np.array_equal(flic.splice_block_rows(parts), whole)


# ---- full-size, size-independent properties (BASELINE.json configs) ----
def _roundtrip_device(codec, imgs, flags=0x01):
    n, h, w, c = imgs.shape
    px = dev(imgs)
    cap = n * int(codec.lib.flic_max_stream_bytes(w, h, c))
    streams = torch.empty(cap, dtype=torch.uint8, device="cuda")
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    out = torch.zeros_like(px)
    codec.encode_batch_device(px, streams, off, flags)
    codec.decode_batch_device(streams, off, out, flags)
    codec.check()
    assert torch.equal(out, px)
    return streams, off.cpu().numpy()


def test_c2_4k_rgba_roundtrip_and_model(codec, oracle):
    import flic_b200 as flic
    img = flic.workloads.make_batch("C2")
    streams, off = _roundtrip_device(codec, img)
    got = streams[: int(off[1])].cpu().numpy()
    assert np.array_equal(got, oracle.encode(img[0]))  # 33 MB through the scalar model: ~0.3 s


def test_c3_batch_slice_roundtrip(codec):
    import flic_b200 as flic
    imgs = flic.workloads.make_batch("C3", n=48)
    streams, off = _roundtrip_device(codec, imgs)
    # images repeat with period 8 -> so must their streams (a checksum-of-checksums style property)
    s = streams.cpu().numpy()
    for i in range(8, 48):
        assert np.array_equal(s[int(off[i]): int(off[i + 1])], s[int(off[i - 8]): int(off[i - 7])])


def test_4k_rgba_alpha_variants(codec, oracle):
    """configs[1] with the three kinds of alpha plane: opaque (flat channel, three symbols per pixel), a
    ramp (four symbols per pixel), and opaque except for one block (both decode paths in one launch)."""
    import flic_b200 as flic
    base = flic.workloads.make_batch("C2")[0]
    ramp = base.copy(); ramp[..., 3] = (np.arange(base.shape[1])[None, :] * 255 // (base.shape[1] - 1)).astype(np.uint8)
    spot = base.copy(); spot[40:50, 200:230, 3] = 7
    imgs = np.stack([base, ramp, spot])
    streams, off = _roundtrip_device(codec, imgs)
    sizes = np.diff(off)
    assert sizes[0] < sizes[2] < sizes[1]
    got = streams[int(off[2]): int(off[3])].cpu().numpy()
    assert np.array_equal(got, oracle.encode(spot))


def test_c5_uniform_noise_roundtrip(codec):
    import flic_b200 as flic
    imgs = flic.workloads.make_batch("C5", n=4)
    _, off = _roundtrip_device(codec, imgs)
    ratio = float(off[-1]) / imgs.size
    assert 1.0 < ratio < 1.03  # incompressible input: 8-bit codes + 1.2 % block headers + 0.8 % slot slack


def test_c4_strip_roundtrip(codec):
    """A 16384-wide, 2048-row strip of C4 (the full 1 GiB image is exercised by bench.py --workload C4)."""
    import flic_b200 as flic
    img = flic.workloads.gradient_noise(16384, 2048, 4, 4)[None]
    _roundtrip_device(codec, img)
