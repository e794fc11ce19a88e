"""A/B of several builds of the engine in ONE process (GPU box): per-kernel CUDA-event times per build and workload,
round trip checked.  usage: python tools/ab.py "libA,libB" "C2x64,C3" [steps]   ("-" = the in-tree libflicb200.so;
other names are relative to the package directory, e.g. exp/base.so)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flic_b200
from bench import WORKLOADS
codec_mod = sys.modules[flic_b200.Codec.__module__]
PKG = os.path.dirname(codec_mod.__file__)
libs = (sys.argv[1] if len(sys.argv) > 1 else "exp/base.so,-").split(",")
wls = (sys.argv[2] if len(sys.argv) > 2 else "C2x64").split(",")
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
flags = int(sys.argv[4], 0) if len(sys.argv) > 4 else 1
st = torch.cuda.current_stream().cuda_stream
for wl in wls:
    cfg, n = WORKLOADS[wl]
    batch = flic_b200.workloads.make_batch(cfg, n=n)
    px = torch.from_numpy(batch).cuda()
    _, h, w, c = batch.shape
    streams = torch.empty(n * flic_b200.max_stream_bytes(w, h, c), dtype=torch.uint8, device="cuda")
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    out = torch.empty_like(px)
    for rep in range(2):
        for lib in libs:
            if lib == "-": os.environ.pop("FLIC_LIB", None)
            else: os.environ["FLIC_LIB"] = os.path.join(PKG, lib)
            codec_mod._lib = None
            codec = flic_b200.Codec(0)
            out.zero_()
            for _ in range(3):
                codec.encode_batch_device(px, streams, off, flags, st); codec.decode_batch_device(streams, off, out, flags, st)
            torch.cuda.synchronize()
            codec.kernel_times(); codec.set_kernel_timing(True)
            for _ in range(steps):
                codec.encode_batch_device(px, streams, off, flags, st); codec.decode_batch_device(streams, off, out, flags, st)
            torch.cuda.synchronize()
            kt = codec.kernel_times()
            res = {k: round(ms / max(cnt, 1), 4) for k, (ms, cnt) in kt.items() if cnt}
            print(json.dumps({"wl": wl, "lib": lib, "rep": rep, "ok": bool(torch.equal(out, px)), "bytes": int(off[-1]), "ms": res}), flush=True)
            codec.close()
    del px, streams, out
