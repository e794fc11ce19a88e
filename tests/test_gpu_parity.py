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


@pytest.fixture(autouse=True)
def _default_encoder_after_each_test(codec):
    yield
    codec.set_encoder("auto")


@pytest.fixture(params=["fused", "staged"])
def encoder(request, codec):
    """Both encode paths (the fused single-pass kernel and the round-1 staged pipeline) must emit the same bytes."""
    codec.set_encoder(request.param)
    yield request.param
    codec.set_encoder("auto")


@pytest.mark.parametrize("name,build", cases.SMALL, ids=[n for n, _ in cases.SMALL])
@pytest.mark.parametrize("flags", [0x01, 0x11])
def test_stage_histograms_and_tables(codec, oracle, encoder, name, build, flags):
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
@pytest.mark.parametrize("flags", cases.ALL_FLAGS)
def test_staged_encoder_bytes(codec, oracle, name, build, flags):
    """The staged pipeline in every layout: slots, ONE_STREAM (k_slots with exact sizes, bit-exact row concatenation in
    k_pack) and EXACT (k_pack places blocks with a look-back over the packed sizes)."""
    codec.set_encoder("staged")
    try:
        img = build()
        assert np.array_equal(codec.encode(img, flags), oracle.encode(img, flags))
    finally:
        codec.set_encoder("auto")


@pytest.mark.parametrize("name,build", cases.SMALL, ids=[n for n, _ in cases.SMALL])
@pytest.mark.parametrize("flags", cases.ALL_FLAGS)
def test_encode_bytes_and_decode_pixels(codec, oracle, name, build, flags):
    codec.set_encoder("fused")
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
        flags |= int(rng.choice([0, 0, 0x20, 0x40]))   # layouts: slots, one stream per block, exact sizes
        want = oracle.encode(img, flags)
        got = codec.encode(img, flags)
        assert np.array_equal(got, want), (trial, w, h, c, flags, kind)
        assert np.array_equal(codec.decode(got), img), (trial, w, h, c, flags, kind)
        assert np.array_equal(oracle.decode(got, img.shape), img), (trial, w, h, c, flags, kind)


@pytest.mark.parametrize("flags", [0x21, 0x41])
def test_corrupt_streams_never_crash_layouts(codec, oracle, flags):
    """The same damage sweep for the two optional layouts.  ONE_STREAM is the delicate one: its decoder follows code
    lengths through the stream with no row structure to bound it, so every loop has to be bounded by validated
    header fields (chains by the block's word count, rounds by the thread count, symbols by the block's pixel
    count) and a damaged code table must not be able to stall a chain (zero-length LUT entries are made to consume
    a bit)."""
    import flic_b200 as flic
    rng = np.random.default_rng(78 + flags)
    img = cases.gradient(384, 96, 4, 45)
    good = codec.encode(img, flags)
    outcomes = {0: 0, -3: 0}
    for trial in range(60):
        bad = good.copy()
        lo = 0 if trial % 3 else 32
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
        assert oracle.decode_rc(bad, img.shape) in (0, -3)
    assert outcomes[0] and outcomes[-3]
    # a code table with holes (an incomplete code): the first block's length nibbles zeroed except two symbols
    nb = int(np.frombuffer(good[20:24], np.uint32)[0])
    first_block = 32 + 4 * (nb + 1) + 4 * int(np.frombuffer(good[32:36], np.uint32)[0])
    bad = good.copy(); bad[first_block: first_block + 128] = 0; bad[first_block] = 0x33   # symbols 0 and 1: 3 bits each
    try:
        codec.decode(bad)
    except flic.FlicError as e:
        assert e.code == -3
    assert np.array_equal(codec.decode(good), img)


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


def test_c3_batch_slice_roundtrip(codec, oracle):
    import flic_b200 as flic
    imgs = flic.workloads.make_batch("C3", n=48)
    streams, off = _roundtrip_device(codec, imgs)
    # one full-size 1080p RGB image (1080 = 33.75 block rows: ragged bottom edge at a 5760-byte pitch) byte for byte
    assert np.array_equal(streams[int(off[5]): int(off[6])].cpu().numpy(), oracle.encode(imgs[5]))
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


def test_c5_uniform_noise_roundtrip(codec, oracle):
    import flic_b200 as flic
    imgs = flic.workloads.make_batch("C5", n=4)
    streams, off = _roundtrip_device(codec, imgs)
    # one full-size noise image (every block at the worst-case code lengths) byte for byte against the model
    assert np.array_equal(streams[int(off[2]): int(off[3])].cpu().numpy(), oracle.encode(imgs[2]))
    ratio = float(off[-1]) / imgs.size
    assert 1.0 < ratio < 1.03  # incompressible input: 8-bit codes + 1.2 % block headers + 0.8 % slot slack


def test_c4_strip_roundtrip(codec):
    """A 16384-wide, 2048-row strip of C4 (the full 1 GiB image is exercised by bench.py --workload C4)."""
    import flic_b200 as flic
    img = flic.workloads.gradient_noise(16384, 2048, 4, 4)[None]
    _roundtrip_device(codec, img)


# ---- round 2: layouts, fused encoder, async host API, device splice ----
@pytest.mark.parametrize("flags", [0x21, 0x41])
def test_layout_modes_full_size(codec, oracle, encoder, flags):
    """ONE_STREAM (self-synchronising decoder) and EXACT (look-back over packed sizes) on full-size inputs, through the
    fused and the staged encoder: a 4K RGBA gradient image, a 4K RGBA noise image (code lengths 7-9: the slowest to
    synchronise) and a 1080p RGB image, each byte-compared with the model and decoded back."""
    import flic_b200 as flic
    for imgs in (flic.workloads.make_batch("C2"), flic.workloads.make_batch("C5", n=1), flic.workloads.make_batch("C3", n=2)):
        streams, off = _roundtrip_device(codec, imgs, flags)
        got = streams[: int(off[1])].cpu().numpy()
        assert np.array_equal(got, oracle.encode(imgs[0], flags))


def test_lookback_state_is_shared_between_the_encoders(codec, oracle):
    """The fused kernel and the staged EXACT packer draw tickets from one counter and tag one status array with one epoch:
    alternate them, launch after launch, on a batch of many blocks (24 images x 3 x 9 blocks) and compare every image."""
    imgs = np.stack([cases.gradient(300, 270, 3, 100 + s) if s % 3 else cases.noise(300, 270, 3, 100 + s) for s in range(24)])
    want = {fl: [oracle.encode(im, fl) for im in imgs] for fl in (0x41, 0x21, 0x01)}
    for rep in range(3):
        for enc in ("staged", "fused"):
            codec.set_encoder(enc)
            for fl in (0x41, 0x21, 0x01):
                streams, off = _roundtrip_device(codec, imgs, fl)
                got = streams[: int(off[-1])].cpu().numpy()
                for i in range(len(imgs)):
                    assert np.array_equal(got[int(off[i]): int(off[i + 1])], want[fl][i]), (rep, enc, hex(fl), i)


def test_one_stream_pathological_codes(codec, oracle):
    """Inputs whose codes synchronise badly or not at all: two symbols with 1-bit codes, 2^k equiprobable symbols
    (fixed-length codes never re-synchronise: the correction has to travel thread by thread), a single symbol."""
    rng = np.random.default_rng(5)
    h, w = 64, 256
    imgs = []
    for nsym in (1, 2, 4, 16, 256):
        steps = rng.integers(0, nsym, size=(h, w, 1)).astype(np.uint8)
        imgs.append(np.cumsum(steps, axis=1).astype(np.uint8))          # residuals uniform over nsym values
    for img in imgs:
        for c in (1, 3):
            im = np.repeat(img, c, axis=2) if c > 1 else img
            want = oracle.encode(im, 0x21)
            got = codec.encode(im, 0x21)
            assert np.array_equal(got, want)
            assert np.array_equal(codec.decode(got), im)


def test_decode_rejects_mismatched_geometry(codec, oracle):
    """a stream of another channel count / colour transform / layout must not decode into plausible
    garbage when the caller passes the wrong geometry to the device API."""
    import flic_b200 as flic
    img = cases.gradient(256, 64, 4, 50)
    s = codec.encode(img, 0x11)
    d_s, d_o = dev(s), torch.tensor([0, s.size], dtype=torch.int64, device="cuda")
    for shape, flags in (((1, 64, 256, 4), 0x01), ((1, 64, 256, 4), 0x31), ((1, 64, 341, 3), 0x11), ((1, 32, 256, 4), 0x11)):
        out = torch.zeros(shape, dtype=torch.uint8, device="cuda")
        codec.decode_batch_device(d_s, d_o, out, flags)
        with pytest.raises(flic.FlicError) as e:
            codec.check()
        assert e.value.code == -3, (shape, flags)
    out = torch.zeros((1, 64, 256, 4), dtype=torch.uint8, device="cuda")
    codec.decode_batch_device(d_s, d_o, out, 0x51)   # EXACT is an encoder-side property: same decode
    codec.check()
    assert np.array_equal(out.cpu().numpy()[0], img)


def test_mixed_geometry_batch(codec, oracle):
    """The host decode API takes batches of mixed geometry (runs of equal geometry decode together)."""
    imgs = [cases.gradient(300, 70, 3, 1), cases.gradient(300, 70, 3, 2), cases.gradient(256, 64, 4, 3),
            cases.noise(129, 65, 3, 4), cases.gradient(300, 70, 3, 5)]
    flags = [0x01, 0x41, 0x21, 0x11, 0x01]   # the first two differ only in EXACT: one run
    streams = [codec.encode(im, f) for im, f in zip(imgs, flags)]
    off = np.concatenate([[0], np.cumsum([s.size for s in streams])]).astype(np.uint64)
    out = np.zeros(sum(im.size for im in imgs), dtype=np.uint8)
    codec.decode_batch(np.concatenate(streams), off, out=out)
    pos = 0
    for im in imgs:
        assert np.array_equal(out[pos: pos + im.size].reshape(im.shape), im)
        pos += im.size


def test_submit_wait_overlap(codec, oracle):
    """flic_encode_submit / flic_decode_submit / flic_wait: one encode and one decode in flight together on one
    context, on pinned buffers; a second submit of the same kind is refused with FLIC_E_BUSY."""
    import flic_b200 as flic
    a = np.stack([cases.gradient(640, 200, 4, 300 + s) for s in range(6)])
    b = np.stack([cases.noise(640, 200, 4, 400 + s) for s in range(6)])
    pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
    a, b = pin(a), pin(b)
    sb, ob = codec.encode_batch(b)
    sb, ob = pin(sb.copy()), ob.copy()
    out_b = pin(np.zeros_like(b))
    cap = a.shape[0] * flic.max_stream_bytes(640, 200, 4)
    out_a, off_a = pin(np.zeros(cap, np.uint8)), np.zeros(a.shape[0] + 1, np.uint64)
    codec.encode_submit(a, out=out_a, offsets=off_a)
    codec.decode_submit(sb, ob, out_b)
    with pytest.raises(flic.FlicError) as e:
        codec.encode_submit(a, out=out_a, offsets=off_a)
    assert e.value.code == -8
    codec.wait(flic.OP_DECODE)
    codec.wait(flic.OP_ENCODE)
    assert np.array_equal(out_b, b)
    want = np.concatenate([oracle.encode(im) for im in a])
    assert off_a[-1] == want.size and np.array_equal(out_a[: want.size], want)
    with pytest.raises(flic.FlicError):
        codec.wait(flic.OP_ENCODE)   # nothing in flight


def test_pageable_and_pinned_host_buffers(codec, oracle, monkeypatch):
    """Pageable caller buffers are pinned for the call (>= 1 MiB) or copied through the driver's staging; both
    give the same bytes as pinned ones."""
    imgs = np.stack([cases.gradient(1024, 300, 4, 500 + s) for s in range(3)])   # 3.7 MB: above the auto-pin threshold
    want = np.concatenate([oracle.encode(im) for im in imgs])
    s1, o1 = codec.encode_batch(imgs)
    assert np.array_equal(s1, want)
    monkeypatch.setenv("FLIC_NO_AUTOPIN", "1")
    s2, o2 = codec.encode_batch(imgs)
    assert np.array_equal(s2, want) and np.array_equal(o1, o2)
    pinned = torch.from_numpy(imgs).pin_memory().numpy()
    s3, _ = codec.encode_batch(pinned)
    assert np.array_equal(s3, want)
    assert np.array_equal(codec.decode_batch(s3, o1), imgs)


@pytest.mark.parametrize("flags", [0x01, 0x21, 0x41])
def test_device_splice_and_split(codec, oracle, flags):
    """The C4 path on one GPU: block-row parts spliced on the device (D2D copies + one kernel) equal the encode of
    the whole image; the plan/finish pair used across GPUs gives the same bytes; and the inverse (a run of block
    rows cut out of the whole stream) decodes to those rows."""
    import flic_b200 as flic
    img = cases.gradient(700, 200, 4, 43)
    whole = codec.encode(img, flags)
    cuts = [(0, 96), (96, 160), (160, 200)]
    parts = [codec.encode(img[a:b], flags) for a, b in cuts]
    d_parts = [dev(p) for p in parts]
    out = torch.zeros(whole.size + 64, dtype=torch.uint8, device="cuda")
    n = codec.splice_block_rows_device(d_parts, out)
    codec.check()
    assert n == whole.size and np.array_equal(out[:n].cpu().numpy(), whole)
    assert np.array_equal(flic.splice_block_rows(parts), whole)
    # plan + finish: what every rank does after the all-gather of (n_blocks, payload_words)
    infos = [flic.peek(p) for p in parts]
    nbs, pws = [i["n_blocks"] for i in infos], [i["payload_words"] for i in infos]
    doff, poff, total = flic.splice_plan(nbs, pws)
    assert total == whole.size
    out2 = torch.zeros(total, dtype=torch.uint8, device="cuda")
    for p, nb, pw, do, po in zip(d_parts, nbs, pws, doff, poff):
        out2[do: do + 4 * nb] = p[32: 32 + 4 * nb]
        out2[po: po + 4 * pw] = p[32 + 4 * (nb + 1): 32 + 4 * (nb + 1) + 4 * pw]
    codec.splice_finish_device(out2, nbs, pws, 700, 200, 4, flags)
    codec.check()
    assert np.array_equal(out2.cpu().numpy(), whole)
    # split: block rows 3..4 (pixel rows 96..160) cut out of the whole stream
    nbx = -(-700 // 128)
    d_whole = dev(whole)
    dirw = np.frombuffer(whole[32: 32 + 4 * (sum(nbs) + 1)].tobytes(), np.uint32)
    b0, b1 = 3 * nbx, 5 * nbx
    w0, w1 = int(dirw[b0]), int(dirw[b1])
    part = torch.zeros(32 + 4 * (b1 - b0 + 1) + 4 * (w1 - w0), dtype=torch.uint8, device="cuda")
    part[32: 32 + 4 * (b1 - b0 + 1)] = d_whole[32 + 4 * b0: 32 + 4 * (b1 + 1)]
    pay0 = 32 + 4 * (sum(nbs) + 1)
    part[32 + 4 * (b1 - b0 + 1):] = d_whole[pay0 + 4 * w0: pay0 + 4 * w1]
    codec.split_finish_device(part, 700, 64, 4, flags)
    codec.check()
    assert np.array_equal(part.cpu().numpy(), parts[1])
    assert np.array_equal(codec.decode(part.cpu().numpy()), img[96:160])


@pytest.mark.parametrize("flags", [0x01, 0x11])
def test_peer_memory_split_on_one_gpu(codec, oracle, flags):
    """The peer-memory form of the block-row split (flic_encode_plan_device / _emit_device / flic_splice_header_device /
    flic_pull_part_device), with one GPU playing every rank in turn and its own memory as the "peer" buffer: the spliced
    stream equals the model's encode of the whole image, and every part pulled back out decodes to its rows.  No value
    ever visits the host between the calls (sizes and bases stay in device memory), as across GPUs."""
    import flic_b200 as flic
    from flic_b200 import sharding
    for img in (cases.gradient(700, 200, 4, 43), cases.noise(300, 131, 3, 44), cases.gradient(130, 33, 1, 45)):
        h, w, c = img.shape
        k = 3
        rows = [sharding.block_row_slice(h, r, k) for r in range(k)]
        nbx = -(-w // 128)
        nbs = [nbx * (-(-(b - a) // 32)) for a, b in rows]
        first = [0] + list(np.cumsum(nbs))
        total_blocks = int(first[-1])
        cap = flic.max_stream_bytes(w, h, c)
        full = torch.full((cap,), 0xA5, dtype=torch.uint8, device="cuda")
        totals = torch.zeros(k, dtype=torch.int64, device="cuda")
        bases = torch.zeros(k + 1, dtype=torch.int64, device="cuda")
        px = dev(img)[None]
        for r, (a, b) in enumerate(rows):
            if b <= a:
                continue
            codec.encode_plan_device(px[:, a:b].contiguous(), flags, totals[r:])
            bases[1:] = torch.cumsum(totals, 0)                  # what the all-gather + prefix sum gives every rank
            codec.encode_emit_device(full.data_ptr(), cap, total_blocks, int(first[r]), bases[r:])
        codec.splice_header_device(full.data_ptr(), cap, w, h, c, flags, bases[k:])
        codec.check()
        want = oracle.encode(img, flags)
        size = 32 + 4 * (total_blocks + 1) + 4 * int(bases[k])
        assert size == want.size
        assert np.array_equal(full[:size].cpu().numpy(), want)
        assert bool((full[size:] == 0xA5).all()), "wrote past the stream"
        for r, (a, b) in enumerate(rows):
            if b <= a:
                continue
            part = torch.zeros(flic.max_stream_bytes(w, b - a, c), dtype=torch.uint8, device="cuda")
            off = torch.zeros(2, dtype=torch.int64, device="cuda")
            codec.pull_part_device(full.data_ptr(), cap, total_blocks, int(first[r]), nbs[r], part, off[1:])
            codec.split_finish_device(part, w, b - a, c, flags)
            out = torch.zeros((1, b - a, w, c), dtype=torch.uint8, device="cuda")
            codec.decode_batch_device(part, off, out, flags)
            codec.check()
            assert np.array_equal(out[0].cpu().numpy(), img[a:b])
            assert np.array_equal(part[: int(off[1])].cpu().numpy(), oracle.encode(img[a:b], flags))


def test_pull_part_refuses_damaged_streams(codec, oracle):
    """flic_pull_part_device treats the stream as input: a wrong magic, a block count that is not the caller's, a payload
    size beyond the buffer, a directory entry beyond the payload or a descending pair are FLIC_E_FORMAT, and nothing is copied."""
    import flic_b200 as flic
    img = cases.gradient(300, 100, 3, 46)
    good = oracle.encode(img)
    nb = int(flic.peek(good)["n_blocks"])
    nbx = -(-300 // 128)
    for word, value in ((0, 0x12345678), (5, nb + 1), (6, 1 << 30), (8 + nbx, 0xFFFFFFF0), (8 + 2 * nbx, 0)):
        bad = good.copy()
        bad.view(np.uint32)[word] = value
        d = dev(bad)
        part = torch.full((good.size,), 0xA5, dtype=torch.uint8, device="cuda")
        off = torch.ones(2, dtype=torch.int64, device="cuda")
        codec.pull_part_device(d.data_ptr(), d.numel(), nb, nbx, nbx, part, off[1:])
        with pytest.raises(flic.FlicError) as ei:
            codec.check()
        assert ei.value.code == -3, (word, ei.value.code)
        assert int(off[1]) == 0 and bool((part == 0xA5).all())
    d = dev(good)
    part = torch.zeros(good.size, dtype=torch.uint8, device="cuda")
    off = torch.zeros(2, dtype=torch.int64, device="cuda")
    codec.pull_part_device(d.data_ptr(), d.numel(), nb, nbx, nbx, part, off[1:])
    codec.split_finish_device(part, 300, 32, 3)
    codec.check()
    assert np.array_equal(part[: int(off[1])].cpu().numpy(), oracle.encode(img[32:64]))


def test_offsets_beyond_4gib(codec):
    """BASELINE config 3's shape at its full size on one GPU — 1024 x 1080p RGB, 6.4 GB of pixels — with half of the
    images noise, so that the streams total more than 4 GiB: u64 stream offsets, u32 directories per image."""
    import flic_b200 as flic
    n, h, w, c = 1024, 1080, 1920, 3
    uniq = torch.cat([dev(flic.workloads.make_batch("C3", n=4)),
                      dev(np.stack([flic.workloads.uniform_noise(w, h, c, 9000 + i) for i in range(4)]))])
    px = uniq.repeat(n // 8, 1, 1, 1)
    cap = n * int(codec.lib.flic_max_stream_bytes(w, h, c))
    streams = torch.empty(cap, dtype=torch.uint8, device="cuda")
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    codec.encode_batch_device(px, streams, off)
    codec.check()
    o = off.cpu().numpy()
    sizes = np.diff(o)
    assert (sizes[8:] == sizes[:-8]).all()
    assert int(o[-1]) > (1 << 32) + (1 << 28)   # the last ~100 streams start beyond the 32-bit byte range
    out = torch.empty_like(px)
    codec.decode_batch_device(streams, off, out)
    codec.check()
    assert torch.equal(out[-8:], uniq) and torch.equal(out[512:520], uniq)
    # the same bytes wherever an image sits in the batch, also past 4 GiB
    for i in (3, 6):
        a, b = int(o[i]), int(o[i + 1])
        a2, b2 = int(o[1016 + i]), int(o[1016 + i + 1])
        assert a2 > (1 << 32) and torch.equal(streams[a:b], streams[a2:b2])
