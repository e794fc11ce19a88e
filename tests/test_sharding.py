"""World-size-2 gloo tests of the multi-GPU host logic (runs on CPU).  The per-rank encode is the
CPU model here — tests may use oracle/ — so what is under test is the partitioning, the tiny
all-gather of byte counts, the gather and the splice, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slices_cover_everything(flic):
    sh = flic.sharding
    for n in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            spans = [sh.batch_slice(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    for h in (1, 31, 32, 33, 2160, 16384):
        for world in (1, 2, 8):
            spans = [sh.block_row_slice(h, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(a % 32 == 0 for a, b in spans if b > a)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, shape, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flic_b200 as flic
        import oracle_binding
        orc = oracle_binding.Oracle(os.path.join(ROOT, "oracle", "libflp0_oracle.so"))
        h, w, c = shape
        img = cases.gradient(w, h, c, 77)
        out = flic.sharding.encode_image_sharded(img, orc.encode, flic.splice_block_rows, dist)
        if rank == 0:
            full = orc.encode(img)
            q.put(("ok", bool(np.array_equal(out, full)), int(out.size)))
        else:
            q.put(("none", out is None, 0))
        # batch path: every rank encodes its slice, no collective; sizes all-gathered only for the check
        lo, hi = flic.sharding.batch_slice(5, rank, world)
        mine = torch.tensor([hi - lo], dtype=torch.int64)
        got = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, mine)
        q.put(("batch", sum(int(t) for t in got) == 5, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(200, 300, 3), (33, 130, 4), (20, 64, 1)])
def test_block_row_split_world2(flic, oracle, shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(4)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _ in res), res


# ---- the device path's host logic (ShardedImageCodec) on CPU tensors over gloo, with the engine replaced by the model
class _ModelCodec:
    """Stand-in for flic_b200.Codec on CPU tensors: exactly the calls ShardedImageCodec makes, answered by the CPU
    model (encode / decode) and by numpy restatements of the two finish kernels."""

    def __init__(self, orc):
        self.orc = orc

    def encode_batch_device(self, px, streams, off, flags, stream=0):
        s = self.orc.encode(px[0].numpy(), flags)
        streams[: s.size] = torch.from_numpy(s)
        off[0], off[1] = 0, int(s.size)

    def decode_batch_device(self, streams, off, out, flags, stream=0):
        n = int(off[1])
        out[0] = torch.from_numpy(self.orc.decode(streams[:n].numpy(), tuple(out.shape[1:])))

    @staticmethod
    def _header(w, h, c, flags, nb, pw):
        return np.array([0x30504C46, 3 | (c << 16) | (flags << 24), w, h, 128 | (32 << 16), nb, pw, 10], dtype=np.uint32)

    def splice_finish_device(self, out, nbs, pws, w, h, c, flags, stream=0):
        o = out.numpy().view(np.uint32)
        nb, pw = sum(nbs), sum(pws)
        first, base = 0, 0
        for n, p in zip(nbs, pws):
            o[8 + first: 8 + first + n] += np.uint32(base)
            first += n; base += p
        o[8 + nb] = pw
        o[:8] = self._header(w, h, c, flags, nb, pw)

    def split_finish_device(self, part, w, h, c, flags, stream=0):
        o = part.numpy().view(np.uint32)
        nb = -(-w // 128) * -(-h // 32)
        first = int(o[8])
        o[8: 8 + nb + 1] -= np.uint32(first)
        o[:8] = self._header(w, h, c, flags, nb, int(o[8 + nb]))


def _worker_dev(rank, world, port, shape, flags, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flic_b200 as flic
        import oracle_binding
        orc = oracle_binding.Oracle(os.path.join(ROOT, "oracle", "libflp0_oracle.so"))
        h, w, c = shape
        img = cases.gradient(w, h, c, 78)
        sc = flic.sharding.ShardedImageCodec(_ModelCodec(orc), w, h, c, flags, dist, rank, world, device="cpu")
        rows = torch.from_numpy(img[sc.y0: sc.y1][None].copy())
        full = sc.encode(rows)
        ok = True
        if rank == 0:
            ok = bool(np.array_equal(full.numpy(), orc.encode(img, flags)))
        out = sc.decode(full)
        ok = ok and bool(np.array_equal(out[0].numpy(), img[sc.y0: sc.y1]))
        q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,flags", [((200, 300, 3), 0x01), ((96, 130, 4), 0x21), ((70, 64, 1), 0x41)])
def test_sharded_codec_exchange_world2(flic, oracle, shape, flags):
    """encode: parts -> all-gather of (n_blocks, payload_words) -> sends straight into the spliced stream -> finish;
    decode: directory cut at block-row boundaries -> sends -> finish -> per-rank rows.  Spliced bytes == the model's
    encode of the whole image; every rank gets its rows back."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_dev, args=(r, 2, port, shape, flags, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(res), res


# ---- the peer-memory path's host logic (PeerImageCodec) over gloo: the "peer buffer" is a file both processes map, the
# engine is replaced by the model and numpy restatements of the small kernels
class _ModelPeerCodec(_ModelCodec):
    def encode_plan_device(self, rows, flags, d_payload_words, stream=0):
        self.planned = self.orc.encode(rows[0].numpy(), flags).view(np.uint32)
        d_payload_words[0] = int(self.planned[6])

    def encode_emit_device(self, root, capacity_bytes, total_blocks, first_block, d_base_words, stream=0):
        o, p = root.numpy().view(np.uint32), self.planned
        nb, pw, base = int(p[5]), int(p[6]), int(d_base_words[0])
        o[8 + first_block: 8 + first_block + nb] = p[8: 8 + nb] + np.uint32(base)
        pay = 8 + total_blocks + 1 + base
        assert 4 * (pay + pw) <= capacity_bytes
        o[pay: pay + pw] = p[8 + nb + 1: 8 + nb + 1 + pw]

    def splice_header_device(self, root, capacity_bytes, w, h, c, flags, d_total_words, stream=0):
        o = root.numpy().view(np.uint32)
        nb, pw = -(-w // 128) * -(-h // 32), int(d_total_words[0])
        o[:8] = self._header(w, h, c, flags, nb, pw)
        o[8 + nb] = pw

    def pull_part_device(self, root, stream_bytes, total_blocks, first_block, nb, part, d_part_bytes, stream=0):
        o, q = root.numpy().view(np.uint32), part.numpy().view(np.uint32)
        assert o[0] == 0x30504C46 and o[5] == total_blocks
        b0, b1 = int(o[8 + first_block]), int(o[8 + first_block + nb])
        q[8: 8 + nb + 1] = o[8 + first_block: 8 + first_block + nb + 1]
        q[8 + nb + 1: 8 + nb + 1 + (b1 - b0)] = o[8 + total_blocks + 1 + b0: 8 + total_blocks + 1 + b1]
        d_part_bytes[0] = 4 * (8 + nb + 1 + (b1 - b0))


def _worker_peer(rank, world, port, shape, flags, path, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flic_b200 as flic
        import oracle_binding
        orc = oracle_binding.Oracle(os.path.join(ROOT, "oracle", "libflp0_oracle.so"))
        h, w, c = shape
        img = cases.gradient(w, h, c, 79)

        def alloc(nbytes):  # one file, mapped shared by every rank: rank 0's "buffer" is everybody's
            t = torch.from_file(path, shared=True, size=nbytes, dtype=torch.uint8)
            return t, t

        sc = flic.sharding.PeerImageCodec(_ModelPeerCodec(orc), w, h, c, flags, dist, rank, world, device="cpu", alloc=alloc)
        rows = torch.from_numpy(img[sc.y0: sc.y1][None].copy())
        ok = True
        for _ in range(2):  # twice: the buffer is reused
            full = sc.encode(rows)
            if rank == 0:
                want = orc.encode(img, flags)
                ok = ok and sc.stream_bytes() == want.size and bool(np.array_equal(full.numpy()[: want.size], want))
            out = sc.decode()
            if sc.y1 > sc.y0:  # (a rank that owns no block row has nothing to compare)
                ok = ok and bool(np.array_equal(out[0].numpy(), img[sc.y0: sc.y1]))
        q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("shape,flags", [((200, 300, 3), 0x01), ((96, 130, 4), 0x11), ((40, 64, 1), 0x01)])
def test_peer_codec_protocol(flic, oracle, tmp_path, shape, flags, world):
    """plan -> all-gather of payload sizes (stays in a tensor) -> prefix sum -> every rank writes its payload and directory
    entries at its base in rank 0's buffer -> fence -> header; decode: fence -> every rank pulls its part -> finish -> rows.
    Spliced bytes == the model's encode of the whole image, also when a rank owns no block row (40 rows over 3 ranks)."""
    import flic_b200 as flic_mod
    h, w, c = shape
    path = str(tmp_path / "peer.bin")
    with open(path, "wb") as f:
        f.truncate(flic_mod.max_stream_bytes(w, h, c))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_peer, args=(r, world, port, shape, flags, path, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(res), res
