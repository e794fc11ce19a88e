"""Two host threads, each pacing its own copies with per-chunk host waits (what flic_encode_submit / flic_decode_submit
do), versus everything enqueued up front (tools/pcie_bw2.py).  No kernels, no library: only the copy pattern."""
import threading, time, torch
MB = 1 << 20
big, small, reps, depth = 33 * MB, 15 * MB, 64, 4
pin = lambda n: torch.empty(n, dtype=torch.uint8).pin_memory()
h_px, h_st, h_st2, h_out = pin(big * reps), pin(small * reps), pin(small * reps), pin(big * reps)
d_e = [torch.empty(big, dtype=torch.uint8, device="cuda") for _ in range(depth)]
d_es = [torch.ones(small, dtype=torch.uint8, device="cuda") for _ in range(depth)]
d_d = [torch.empty(small, dtype=torch.uint8, device="cuda") for _ in range(depth)]
d_dp = [torch.ones(big, dtype=torch.uint8, device="cuda") for _ in range(depth)]
sA, sB, sC, sD = (torch.cuda.Stream() for _ in range(4))
def pipe(h_src, d_in, d_res, h_dst, s_in, s_out, n_in, n_out):
    ev = [torch.cuda.Event(blocking=True) for _ in range(depth)]
    def issue(k):
        b = k % depth
        with torch.cuda.stream(s_in):
            d_in[b].copy_(h_src[k * n_in:(k + 1) * n_in], non_blocking=True)
            ev[b].record(s_in)
    issued = 0
    while issued < min(depth, reps): issue(issued); issued += 1
    for k in range(reps):
        b = k % depth
        ev[b].synchronize()                     # "the kernel of chunk k is done"
        with torch.cuda.stream(s_out):
            h_dst[k * n_out:(k + 1) * n_out].copy_(d_res[b], non_blocking=True)
        if issued < reps: issue(issued); issued += 1
    s_in.synchronize(); s_out.synchronize()
enc = lambda: pipe(h_px, d_e, d_es, h_st2, sA, sB, big, small)
dec = lambda: pipe(h_st, d_d, d_dp, h_out, sC, sD, small, big)
def T(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0)
def both():
    t = threading.Thread(target=dec); t.start(); enc(); t.join()
enc(); dec(); both()
for name, f in (("encode-like alone", enc), ("decode-like alone", dec), ("both, two host threads", both), ("both, two host threads", both)):
    print(f"{name:26s} {T(f):7.1f} ms")

# variant: the decode-like side is enqueued up front with device-side dependencies only (no host pacing)
def dec_pre():
    evs = []
    for k in range(reps):
        b = k % depth
        with torch.cuda.stream(sC):
            if k >= depth: sC.wait_event(evs[k - depth])     # its staging buffer was read by D2H(k - depth)
            d_d[b].copy_(h_st[k * small:(k + 1) * small], non_blocking=True)
            e = torch.cuda.Event(); e.record(sC)
        with torch.cuda.stream(sD):
            sD.wait_event(e)
            h_out[k * big:(k + 1) * big].copy_(d_dp[b], non_blocking=True)
            e2 = torch.cuda.Event(); e2.record(sD); evs.append(e2)
    sC.synchronize(); sD.synchronize()
def both_pre():
    t = threading.Thread(target=dec_pre); t.start(); enc(); t.join()
dec_pre(); both_pre()
for name, f in (("decode-like pre-enqueued", dec_pre), ("both, decode pre-enqueued", both_pre), ("both, decode pre-enqueued", both_pre)):
    print(f"{name:26s} {T(f):7.1f} ms")
for dpt in (8, 16):
    depth = dpt
    d_e = [torch.empty(big, dtype=torch.uint8, device="cuda") for _ in range(depth)]
    d_es = [torch.ones(small, dtype=torch.uint8, device="cuda") for _ in range(depth)]
    d_d = [torch.empty(small, dtype=torch.uint8, device="cuda") for _ in range(depth)]
    d_dp = [torch.ones(big, dtype=torch.uint8, device="cuda") for _ in range(depth)]
    enc = lambda: pipe(h_px, d_e, d_es, h_st2, sA, sB, big, small)
    dec = lambda: pipe(h_st, d_d, d_dp, h_out, sC, sD, small, big)
    both(); both_pre()
    print(f"depth {dpt}: both host-paced {T(both):7.1f} ms; decode pre-enqueued {T(both_pre):7.1f} ms")
