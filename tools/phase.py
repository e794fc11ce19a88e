"""Per-phase cycle breakdown of the fused encoder (FLIC_PHASE_CLOCKS=1): python tools/phase.py WORKLOAD [flags]"""
import os, sys
os.environ["FLIC_PHASE_CLOCKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, flic_b200 as flic
sys.path.insert(0, ROOT)
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "C2x8"
flags = int(sys.argv[2], 0) if len(sys.argv) > 2 else 1
cfg, n = bench.WORKLOADS[wl]
batch = flic.workloads.make_batch(cfg, n=n)
px = torch.from_numpy(batch).cuda()
_, h, w, c = batch.shape
codec = flic.Codec(0)
codec.set_encoder('fused')
streams = torch.empty(n * flic.max_stream_bytes(w, h, c), dtype=torch.uint8, device="cuda")
off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
for _ in range(3):
    codec.encode_batch_device(px, streams, off, flags)
out = torch.empty_like(px)
codec.check(); codec.phase_clocks()
K = 5
for _ in range(K):
    codec.encode_batch_device(px, streams, off, flags)
    if flags & 0x20:
        codec.decode_batch_device(streams, off, out, flags)
codec.check()
cyc_all = codec.phase_clocks()
cyc = cyc_all[:8]
nblk = n * (-(-w // 128)) * (-(-h // 32)) * K
names = ["ticket+clear", "load+resid+hist", "hist reduce", "table", "pack", "look-back", "copy-out", "-"]
tot = sum(cyc)
print(wl, hex(flags), "blocks", nblk, "cycles/block", round(tot / nblk))
for nme, cy in zip(names, cyc):
    print(f"  {nme:18s} {cy / nblk:9.0f} cyc/block  {100 * cy / max(tot, 1):5.1f} %")

if flags & 0x20:
    cyc = cyc_all[8:]
    names = ["stream copy + LUT", "speculative chains", "correction rounds", "2b offset table", "2b walk", "scan", "final decode", "un-predict + store"]
    tot = sum(cyc)
    print("k_decode_one: cycles/block", round(tot / nblk))
    for nme, cy in zip(names, cyc):
        print(f"  {nme:18s} {cy / nblk:9.0f} cyc/block  {100 * cy / max(tot, 1):5.1f} %")
