"""Is one encode + one decode in flight together faster than one after the other?  (host-buffer API, pinned buffers)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, flic_b200 as flic
n = 64
batch = flic.workloads.make_batch("C2", n=n)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
h_in = pin(batch)
cap = n * flic.max_stream_bytes(3840, 2160, 4)
h_str = pin(np.zeros(cap, np.uint8)); h_off = np.zeros(n + 1, np.uint64)
h_str2 = pin(np.zeros(cap, np.uint8)); h_off2 = np.zeros(n + 1, np.uint64)
h_out = pin(np.zeros_like(batch))
codec = flic.Codec(0)
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0)
def enc(buf=h_str, off=h_off):
    codec.encode_submit(h_in, out=buf, offsets=off); codec.wait(flic.OP_ENCODE)
def dec():
    nb = int(h_off[n]); codec.decode_submit(h_str[:nb], h_off, h_out); codec.wait(flic.OP_DECODE)
def both():
    nb = int(h_off[n])
    codec.encode_submit(h_in, out=h_str2, offsets=h_off2)
    codec.decode_submit(h_str[:nb], h_off, h_out)
    codec.wait(flic.OP_DECODE); codec.wait(flic.OP_ENCODE)
enc(); dec(); both()
for name, f in (("encode alone", enc), ("decode alone", dec), ("both together", both), ("encode alone", enc), ("decode alone", dec), ("both together", both)):
    print(f"{name:14s} {T(f):7.1f} ms")
# two contexts, two Python threads, blocking calls: does THAT overlap?
import threading
codec2 = flic.Codec(0)
def both2():
    nb = int(h_off[n])
    t = threading.Thread(target=lambda: codec2.decode_batch(h_str[:nb], h_off, out=h_out))
    t.start(); codec.encode_batch(h_in, out=h_str2, offsets=h_off2); t.join()
both2()
print(f"two contexts   {T(both2):7.1f} ms")
print("raw", batch.nbytes / 1e9, "GB; comp", int(h_off[n]) / 1e9, "GB")
