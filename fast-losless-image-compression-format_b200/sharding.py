"""Multi-GPU partitioning for the two ways the path shards (BASELINE.json north_star):

* a batch of independent images: contiguous slices per rank, no data-path collective;
* one oversized image: whole block rows per rank, then ONE tiny all-gather of the
  per-rank stream byte counts, after which every rank knows where its payload lands
  and rank 0 splices (directory rebase + payload concatenation).

Works on any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
The encode function is injected so the host logic is testable without a GPU.
"""
import numpy as np

from .codec import BLOCK_H


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
