"""Does the link carry H2D and D2H together as well with four copy streams (two per direction, mixed sizes: what one
encode call and one decode call in flight issue) as with one stream per direction?"""
import torch, json
MB = 1 << 20
big, small, reps = 33 * MB, 15 * MB, 64
pin = lambda n: torch.empty(n, dtype=torch.uint8).pin_memory()
hA, hB, hC, hD = pin(big * 4), pin(small * 4), pin(small * 4), pin(big * 4)
dA, dB, dC, dD = (torch.ones(x.numel(), dtype=torch.uint8, device="cuda") for x in (hA, hB, hC, hD))
S = [torch.cuda.Stream() for _ in range(4)]
def run(plan):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in S: s.wait_event(a)
    for i in range(reps):
        o = (i % 4)
        for st, kind, h, d, n in plan:
            with torch.cuda.stream(S[st]):
                if kind == "h2d": d[o * n:(o + 1) * n].copy_(h[o * n:(o + 1) * n], non_blocking=True)
                else: h[o * n:(o + 1) * n].copy_(d[o * n:(o + 1) * n], non_blocking=True)
    for s in S:
        e = torch.cuda.Event(); e.record(s); torch.cuda.current_stream().wait_event(e)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
four = [(0, "h2d", hA, dA, big), (1, "d2h", hB, dB, small), (2, "h2d", hC, dC, small), (3, "d2h", hD, dD, big)]
two = [(0, "h2d", hA, dA, big), (1, "d2h", hB, dB, small), (0, "h2d", hC, dC, small), (1, "d2h", hD, dD, big)]
enc_only = [(0, "h2d", hA, dA, big), (1, "d2h", hB, dB, small)]
dec_only = [(2, "h2d", hC, dC, small), (3, "d2h", hD, dD, big)]
for name, plan in (("warm", four), ("encode-like alone", enc_only), ("decode-like alone", dec_only), ("both, four streams", four), ("both, one stream per direction", two)):
    ms = run(plan)
    gb = sum(n for *_, n in plan) * reps / 1e9
    print(f"{name:32s} {ms:7.1f} ms  aggregate {gb / (ms * 1e-3):6.1f} GB/s")
p = torch.cuda.get_device_properties(0)
print("asyncEngineCount-ish:", getattr(p, "async_engine_count", "n/a"), p.name)
