"""Experiment helper (GPU box): per-kernel CUDA-event times of encode+decode on one workload, no correctness
checks (use tools/quick.sh for parity).  usage: python tools/ktime.py [workload] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flic_b200
from bench import WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "C2x64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg, n = WORKLOADS[wl]
flic_b200.build_library()
codec = flic_b200.Codec(0)
batch = flic_b200.workloads.make_batch(cfg, n=n)
px = torch.from_numpy(batch).cuda()
_, h, w, c = batch.shape
streams = torch.empty(n * flic_b200.max_stream_bytes(w, h, c), dtype=torch.uint8, device="cuda")
off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
out = torch.empty_like(px)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    codec.encode_batch_device(px, streams, off, 1, st); codec.decode_batch_device(streams, off, out, 1, st)
torch.cuda.synchronize()
codec.kernel_times(); codec.set_kernel_timing(True)
for _ in range(steps):
    codec.encode_batch_device(px, streams, off, 1, st); codec.decode_batch_device(streams, off, out, 1, st)
torch.cuda.synchronize()
kt = codec.kernel_times()
print(wl, os.environ.get("FLIC_EXP", ""), {k: round(ms / max(cnt, 1), 4) for k, (ms, cnt) in kt.items()}, "ok" if torch.equal(out, px) else "MISMATCH(expected under FLIC_EXP)")
