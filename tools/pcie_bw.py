"""Host<->device copy bandwidth of the box the bench runs on: H2D alone, D2H alone, both at once (pinned memory).
The e2e number of bench.py cannot exceed raw_bytes / ((raw + compressed) / link_GBps_per_direction)."""
import json, sys, torch
def probe(mb=1024, reps=5):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a); s2.wait_event(a)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s1); e2.record(s2)
        torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
        b.record(); torch.cuda.synchronize()
        return n * reps / (a.elapsed_time(b) * 1e-3) / 1e9
    run(True, True)
    return {"h2d_alone_GBps": round(run(True, False), 1), "d2h_alone_GBps": round(run(False, True), 1),
            "each_direction_when_both_GBps": round(run(True, True), 1)}
if __name__ == "__main__":
    print(json.dumps(probe()))
