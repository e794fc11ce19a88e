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
