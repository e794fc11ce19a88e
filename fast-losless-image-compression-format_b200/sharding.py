"""Multi-GPU partitioning for the two ways the path shards (BASELINE.json north_star):

* a batch of independent images: contiguous slices per rank, no data-path collective;
* one oversized image: whole block rows per rank, then ONE tiny all-gather of the per-rank part sizes,
  after which every rank knows where its directory entries and payload land in the spliced stream and
  sends them straight there; rank 0 finishes the splice (header + directory rebase).

Three implementations of the second path:

* `encode_image_sharded` — host logic over any torch.distributed backend (gloo in the CPU tests), with
  the encode and splice functions injected so it is testable without a GPU;
* `ShardedImageCodec` — a device path: parts never leave HBM, the payloads travel by NCCL send/recv
  (NVLink) directly into their final position, and the directory rebase runs as a kernel
  (`flic_splice_finish_device` / `flic_split_finish_device`).  The sizes visit the host (NCCL needs them
  to address its sends): two host synchronisations per direction;
* `PeerImageCodec` — the path bench.py's C4 line uses when peer memory is available: the spliced stream
  lives in a torch symmetric-memory buffer on rank 0 that every rank has mapped; every rank's pack
  kernel writes its block payloads straight into it over NVLink while it packs (compute and transfer in
  ONE kernel, flic_encode_emit_device), after a device-side all-gather of the part sizes; for decode every
  rank pulls its part out of rank 0's memory with a copy kernel.  Nothing synchronises with the host.
"""
import numpy as np

from .codec import BLOCK_H, BLOCK_W, splice_plan

HEADER_BYTES = 32


def batch_slice(n_images, rank, world):
    """Contiguous [lo, hi) of the batch owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def block_row_slice(height, rank, world):
    """Pixel-row range [y0, y1) of whole block rows owned by `rank` (last rank takes the ragged tail)."""
    nby = (height + BLOCK_H - 1) // BLOCK_H
    lo, hi = batch_slice(nby, rank, world)
    return min(lo * BLOCK_H, height), min(hi * BLOCK_H, height)


def encode_image_sharded(image, encode_fn, splice_fn, dist=None, device="cpu"):
    """Encode one image split by block rows across the ranks of `dist`.

    encode_fn(sub_image) -> uint8 stream; splice_fn(list_of_streams) -> uint8 stream.
    Returns the full stream on rank 0 and None elsewhere.  Collectives: one all_gather of
    a single int64 per rank (the byte counts), then the variable-size gather of payloads.
    """
    import torch

    if dist is None or not dist.is_initialized():
        return encode_fn(image)
    rank, world = dist.get_rank(), dist.get_world_size()
    y0, y1 = block_row_slice(image.shape[0], rank, world)
    part = encode_fn(image[y0:y1]) if y1 > y0 else np.zeros(0, dtype=np.uint8)

    mine = torch.tensor([part.size], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, mine)  # the "tiny all-gather of per-GPU byte counts"
    counts = [int(t.item()) for t in counts]

    # variable-size gather to rank 0, padded to the largest part
    pad = max(max(counts), 1)
    buf = torch.zeros(pad, dtype=torch.uint8, device=device)
    if part.size:
        buf[: part.size] = torch.from_numpy(np.ascontiguousarray(part)).to(device)
    gathered = [torch.zeros(pad, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    parts = [gathered[r][: counts[r]].cpu().numpy() for r in range(world) if counts[r]]
    return splice_fn(parts)


class ShardedImageCodec:
    """One w x h x c image split by block rows over the ranks of a NCCL process group; everything stays in HBM.

    encode(rows): this rank's pixel rows [1, rows, w, c] (CUDA uint8) -> the spliced stream of the whole image
        on rank 0 (a view into a preallocated buffer; None elsewhere).  Steps: the engine encodes the part;
        the (n_blocks, payload_words) pairs are all-gathered (the only collective: 16 bytes per rank); every
        rank sends its directory entries and its payload to rank 0, which receives them at the positions
        flic_splice_plan gives — no staging copy; rank 0 runs flic_splice_finish_device.
    decode(stream): the inverse — rank 0 cuts the directory at block-row boundaries and sends each rank its
        entries and payload; every rank finishes its part stream (flic_split_finish_device) and decodes its
        rows into its own [1, rows, w, c] tensor.
    Every rank must call both with the same geometry.  world == 1 degenerates to plain encode / decode.
    """

    def __init__(self, codec, w, h, c, flags, dist=None, rank=0, world=1, device="cuda"):
        import torch

        self.torch, self.codec, self.dist = torch, codec, dist
        self.w, self.h, self.c, self.flags, self.rank, self.world = w, h, c, flags, rank, world
        self.rows = [block_row_slice(h, r, world) for r in range(world)]
        self.y0, self.y1 = self.rows[rank]
        self.nbx = (w + BLOCK_W - 1) // BLOCK_W
        self.nbs = [self.nbx * ((b - a + BLOCK_H - 1) // BLOCK_H) for a, b in self.rows]
        my_rows = max(self.y1 - self.y0, 1)
        from .codec import max_stream_bytes
        self.part = torch.empty(max_stream_bytes(w, my_rows, c), dtype=torch.uint8, device=device)
        self.part_off = torch.zeros(2, dtype=torch.int64, device=device)
        self.full = torch.empty(max_stream_bytes(w, h, c), dtype=torch.uint8, device=device) if rank == 0 else None
        self.mine = torch.zeros(2, dtype=torch.int64, device=device)
        self.counts = torch.zeros((world, 2), dtype=torch.int64, device=device)
        self.pws = None        # payload words of every part (set by encode, or by decode on rank 0 and broadcast)
        self.out = torch.empty((1, my_rows, w, c), dtype=torch.uint8, device=device)

    def _exchange(self, ops):
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()

    def encode(self, rows, stream=0):
        torch, dist = self.torch, self.dist
        self.codec.encode_batch_device(rows, self.part, self.part_off, self.flags, stream)
        if self.world == 1:
            n = int(self.part_off[1])
            self.pws = [(n - HEADER_BYTES - 4 * (self.nbs[0] + 1)) // 4]
            return self.part[:n]
        hdr = self.part[:HEADER_BYTES].view(torch.int32)
        self.mine.copy_(hdr[5:7])                         # (n_blocks, payload_words) of this part, on the device
        self.mine.bitwise_and_(0xFFFFFFFF)                # the header words are unsigned
        dist.all_gather_into_tensor(self.counts.view(-1), self.mine)   # the tiny all-gather
        counts = self.counts.cpu().tolist()               # every rank needs the sizes to address its sends
        nbs, pws = [int(x[0]) for x in counts], [int(x[1]) for x in counts]
        assert nbs == self.nbs, (nbs, self.nbs)
        self.pws = pws
        doff, poff, total = splice_plan(nbs, pws)
        r = self.rank
        my_dir = self.part[HEADER_BYTES: HEADER_BYTES + 4 * nbs[r]]
        my_pay = self.part[HEADER_BYTES + 4 * (nbs[r] + 1): HEADER_BYTES + 4 * (nbs[r] + 1) + 4 * pws[r]]
        ops = []
        if r == 0:
            self.full[doff[0]: doff[0] + 4 * nbs[0]].copy_(my_dir)
            self.full[poff[0]: poff[0] + 4 * pws[0]].copy_(my_pay)
            for q in range(1, self.world):
                ops.append(dist.P2POp(dist.irecv, self.full[doff[q]: doff[q] + 4 * nbs[q]], q))
                if pws[q]:
                    ops.append(dist.P2POp(dist.irecv, self.full[poff[q]: poff[q] + 4 * pws[q]], q))
        else:
            ops.append(dist.P2POp(dist.isend, my_dir, 0))
            if pws[r]:
                ops.append(dist.P2POp(dist.isend, my_pay, 0))
        self._exchange(ops)
        if r != 0:
            return None
        self.codec.splice_finish_device(self.full, nbs, pws, self.w, self.h, self.c, self.flags, stream)
        return self.full[:total]

    def decode(self, stream_full=None, stream=0):
        torch, dist = self.torch, self.dist
        if self.world == 1:
            off = torch.tensor([0, stream_full.numel()], dtype=torch.int64, device=stream_full.device)
            self.codec.decode_batch_device(stream_full, off, self.out, self.flags, stream)
            return self.out
        nbs = self.nbs
        k = self.world
        first = np.concatenate([[0], np.cumsum(nbs)]).astype(np.int64)   # first block of every part
        # rank 0 reads the directory at the k+1 part boundaries and tells everybody (the tiny collective of decode)
        bounds = torch.zeros(k + 1, dtype=torch.int64, device=self.part.device)
        if self.rank == 0:
            dirw = stream_full[HEADER_BYTES: HEADER_BYTES + 4 * (int(first[-1]) + 1)].view(torch.int32)
            idx = torch.from_numpy(first).to(self.part.device)
            bounds.copy_(dirw[idx].to(torch.int64) & 0xFFFFFFFF)
        dist.broadcast(bounds, 0)
        b = bounds.cpu().tolist()
        pws = [int(b[i + 1] - b[i]) for i in range(k)]
        pay0 = HEADER_BYTES + 4 * (int(first[-1]) + 1)
        r = self.rank
        my_dir = self.part[HEADER_BYTES: HEADER_BYTES + 4 * (nbs[r] + 1)]
        my_pay = self.part[HEADER_BYTES + 4 * (nbs[r] + 1): HEADER_BYTES + 4 * (nbs[r] + 1) + 4 * pws[r]]
        ops = []
        if r == 0:
            for q in range(1, k):
                ops.append(dist.P2POp(dist.isend, stream_full[HEADER_BYTES + 4 * int(first[q]): HEADER_BYTES + 4 * (int(first[q + 1]) + 1)], q))
                if pws[q]:
                    ops.append(dist.P2POp(dist.isend, stream_full[pay0 + 4 * int(b[q]): pay0 + 4 * int(b[q + 1])], q))
            my_dir.copy_(stream_full[HEADER_BYTES: HEADER_BYTES + 4 * (nbs[0] + 1)])
            my_pay.copy_(stream_full[pay0: pay0 + 4 * pws[0]])
        else:
            ops.append(dist.P2POp(dist.irecv, my_dir, 0))
            if pws[r]:
                ops.append(dist.P2POp(dist.irecv, my_pay, 0))
        self._exchange(ops)
        self.codec.split_finish_device(self.part, self.w, self.y1 - self.y0, self.c, self.flags, stream)
        self.part_off[1] = HEADER_BYTES + 4 * (nbs[r] + 1) + 4 * pws[r]
        self.codec.decode_batch_device(self.part, self.part_off, self.out, self.flags, stream)
        return self.out


class PeerImageCodec:
    """One w x h x c image split by block rows over the ranks of a NCCL process group, spliced through PEER MEMORY.

    The stream buffer is allocated as torch symmetric memory (one buffer per rank, all mapped into every process; only
    rank 0's is written).  encode(rows): every rank plans its part (histograms, tables, slot positions), the payload
    sizes are all-gathered on the device (8 bytes per rank — the only collective that carries data), an exclusive prefix
    sum gives every rank its base, and flic_encode_emit_device packs the part straight into rank 0's buffer over NVLink;
    a one-element all-reduce orders rank 0's header write (and any reader) behind everybody's stores.  decode(): every
    rank copies its directory entries and payload out of rank 0's buffer (flic_pull_part_device), finishes them into a
    stand-alone stream and decodes its rows.  No call synchronises with the host; sizes never leave device memory.
    Raises RuntimeError when symmetric memory cannot be set up (the caller then uses ShardedImageCodec)."""

    def __init__(self, codec, w, h, c, flags, dist, rank, world, device="cuda", alloc=None):
        """alloc (tests): nbytes -> (this rank's buffer, what to hand the codec as rank 0's buffer); the default allocates
        torch symmetric memory and hands the codec rank 0's device address."""
        import torch
        from .codec import max_stream_bytes

        if flags & 0x60:
            raise RuntimeError("the peer-memory split needs the default layout (positions from the histograms)")
        self.torch, self.codec, self.dist = torch, codec, dist
        self.w, self.h, self.c, self.flags, self.rank, self.world = w, h, c, flags, rank, world
        self.rows = [block_row_slice(h, r, world) for r in range(world)]
        self.y0, self.y1 = self.rows[rank]
        self.nbx = (w + BLOCK_W - 1) // BLOCK_W
        self.nbs = [self.nbx * ((b - a + BLOCK_H - 1) // BLOCK_H) for a, b in self.rows]
        self.first = [0]
        for nb in self.nbs:
            self.first.append(self.first[-1] + nb)
        self.total_blocks = self.first[-1]
        my_rows = max(self.y1 - self.y0, 1)
        self.cap = max_stream_bytes(w, h, c)
        if alloc is None:
            # symmetric memory: same size on every rank (only rank 0's buffer holds the stream)
            import torch.distributed._symmetric_memory as symm_mem
            self.full = symm_mem.empty(self.cap, dtype=torch.uint8, device=device)
            self.hdl = symm_mem.rendezvous(self.full, dist.group.WORLD.group_name)
            self.root_ptr = int(self.hdl.buffer_ptrs[0])
        else:
            self.full, self.root_ptr = alloc(self.cap)
        self.part = torch.empty(max_stream_bytes(w, my_rows, c), dtype=torch.uint8, device=device)
        self.part_off = torch.zeros(2, dtype=torch.int64, device=device)
        self.mine = torch.zeros(1, dtype=torch.int64, device=device)
        self.totals = torch.zeros(world, dtype=torch.int64, device=device)
        self.bases = torch.zeros(world + 1, dtype=torch.int64, device=device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.out = torch.empty((1, my_rows, w, c), dtype=torch.uint8, device=device)

    def _fence(self):
        """Orders everything the ranks have queued so far before everything queued after it, on the device: a one-element
        all-reduce runs behind each rank's earlier kernels and completes only when every rank has contributed."""
        self.dist.all_reduce(self.flag)

    def encode(self, rows, stream=0):
        torch, dist, r = self.torch, self.dist, self.rank
        if self.y1 > self.y0:
            self.codec.encode_plan_device(rows, self.flags, self.mine, stream)
        else:
            self.mine.zero_()
        dist.all_gather_into_tensor(self.totals, self.mine)          # the tiny all-gather (sizes stay on the device)
        self.bases[1:] = torch.cumsum(self.totals, 0)
        if self.y1 > self.y0:
            self.codec.encode_emit_device(self.root_ptr, self.cap, self.total_blocks, self.first[r], self.bases[r:], stream)
        self._fence()
        if r == 0:
            self.codec.splice_header_device(self.root_ptr, self.cap, self.w, self.h, self.c, self.flags, self.bases[self.world:], stream)
        return self.full if r == 0 else None

    def stream_bytes(self):
        """Size of the spliced stream (a host read: call it outside timed regions)."""
        return HEADER_BYTES + 4 * (self.total_blocks + 1) + 4 * int(self.bases[self.world].item())

    def decode(self, stream_full=None, stream=0):
        r = self.rank
        self._fence()                                                # rank 0's header is in place before anybody reads
        if self.y1 > self.y0:
            self.codec.pull_part_device(self.root_ptr, self.cap, self.total_blocks, self.first[r], self.nbs[r], self.part, self.part_off[1:], stream)
            self.codec.split_finish_device(self.part, self.w, self.y1 - self.y0, self.c, self.flags, stream)
            self.codec.decode_batch_device(self.part, self.part_off, self.out, self.flags, stream)
        self._fence()                                                # every rank has its part: the buffer may be reused
        return self.out
