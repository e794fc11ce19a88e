#!/usr/bin/env python
"""bench.py — raw-pixel GB/s of the encode+decode hot path (BASELINE.json `metric`).

A step = one encode pass + one decode pass over one batch of synthetic images.
  value     : whole-job raw-pixel GB/s, inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e       : the same metric through the host-buffer C-ABI calls on PINNED host buffers (H2D and D2H inside
              the timed region); encode of step k+1 and decode of step k are in flight together
              (flic_encode_submit / flic_decode_submit / flic_wait), so both directions of the link are busy.
  roofline  : frac = min(encode path, decode path) of (1 + r) * N bytes per pass against the measured HBM copy
              bandwidth in MEASURED_PEAKS.json; per-kernel CUDA-event durations from the library's timing hook.
  secondary : the other BASELINE.json configs, driver-timed in the same run: C3 (1024 x 1080p RGB, STRONG scaling:
              the fixed batch is split over the ranks), C5 (4K RGBA uniform noise, weak), C4 (ONE 16384^2 RGBA image
              split by block rows over the ranks, all-gather + NVLink gather + splice INSIDE the timed region), and
              the two pessimistic stream layouts on the main workload (ONE_STREAM, EXACT).
  cpu_model : the scalar CPU model of the same provisional format on the host cores.

LICENSING GATE: the reference may not be built, run or restated (LICENSING.md), and there is no Rust toolchain in
the image, so the reference's own CPU implementation cannot be timed: `cpu_baseline` says "unavailable" and
`--impl reference` prints {"impl": "reference", "unavailable": ...}.  The CPU number this script does measure is
the scalar C model of the provisional FLP0 format (oracle/), reported as `cpu_model` — it is NOT the reference.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "raw_pixel_GBps_encode_plus_decode"
GATE = ("licensing gate: reference README reserves all use, no licence file; BASELINE.json north_star forbids "
        "building/running/restating it until cleared (LICENSING.md); no Rust toolchain in the image either")
FORMAT = "FLP0 v3 (provisional; NOT the reference bitstream — licensing gate)"
DISTINCT = 8  # images synthesised per workload; the rest of a batch cycles through them (8 x 33 MB still exceeds L2)

# workload name -> (config id, images per GPU per step)
WORKLOADS = {
    "C2x64": ("C2", 64),  # configs[1] geometry (3840x2160 RGBA8) batched so a step exceeds L2 (2.1 GB raw)
    "C2x8": ("C2", 8),    # same geometry, 265 MB/step: short enough to profile under ncu, still > L2
    "C2Ax64": ("C2A", 64),  # the same batch with a gradient alpha plane (no flat channel: four symbols per pixel)
    "C1": ("C1", 1), "C2": ("C2", 1), "C3": ("C3", 128), "C4": ("C4", 1), "C5": ("C5", 64), "C5x8": ("C5", 8),
    "C3full": ("C3", 1024),  # configs[2] at its full size on ONE GPU: 6.4 GB raw per step
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_model_roundtrip(batch, seconds=20.0):
    """Scalar CPU model (oracle/) encode+decode on all host cores over a bounded sample of the batch."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    import oracle_binding
    so = os.path.join(ROOT, "oracle", "libflp0_oracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = oracle_binding.Oracle(so)
    cores = os.cpu_count() or 1
    per_img = batch[0].nbytes
    # ~80 MB/s round trip per core: size the sample to about `seconds` of work
    k = int(max(cores, min(100000, seconds * 80e6 * cores / per_img)))
    idx = [i % len(batch) for i in range(k)]

    def one(i):
        s = orc.encode(batch[i])
        out = orc.decode(s, batch[i].shape)
        return s.size, bool(np.array_equal(out, batch[i]))

    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:  # ctypes releases the GIL
        res = list(ex.map(one, idx))
    dt = time.perf_counter() - t0
    assert all(ok for _, ok in res)
    return {"value": k * per_img / dt / 1e9, "unit": "GB/s", "cores": min(cores, k), "kind": "flp0-model",
            "sample": f"{k} images of {batch.shape[2]}x{batch.shape[1]}x{batch.shape[3]} encode+decode, "
                      f"{min(cores, k)} threads, {dt:.1f} s; scalar C model of FLP0 (oracle/), NOT the gated reference"}


def bind_to_gpu_numa(index):
    """Pin this process's threads to the CPUs of the NUMA node the GPU hangs off, BEFORE the pinned buffers are
    allocated (first touch then places them on that node), so H2D/D2H do not cross the socket interconnect.
    Returns a short description; a no-op where the box exposes one node or hides the topology."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(os.path.join(path, "numa_node")).read())
        cpus = open(os.path.join(path, "local_cpulist")).read().strip()
        allowed = os.sched_getaffinity(0)
        want = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            want.update(range(int(a), int(b or a) + 1))
        want &= allowed
        if node < 0 or not want or want == allowed:
            return f"numa node {node}: no binding needed", None
        os.sched_setaffinity(0, want)
        return f"bound to numa node {node} ({len(want)} cpus)", allowed
    except Exception as e:  # containers often hide /sys topology
        return f"not bound ({type(e).__name__})", None


class Dist:
    """Thin wrapper so single-GPU runs need no process group."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def device_roundtrip(D, codec, px, flags, steps, warmup, kernel_times=False):
    """Encode + decode of the CUDA tensor px [n,h,w,c], inputs resident in HBM.  Per-step CUDA events on the
    launching stream; the L2 is flushed between steps when a step's pixels would fit in it.  Returns a dict with
    max-over-ranks encode/decode ms per step, this rank's compressed bytes, kernel times and launch count."""
    torch = D.torch
    n, h, w, c = px.shape
    raw = px.numel()
    import flic_b200
    cap = n * flic_b200.max_stream_bytes(w, h, c)
    streams = torch.empty(cap, dtype=torch.uint8, device="cuda")
    off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    out = torch.empty_like(px)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if raw < (512 << 20) else None
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(warmup):
        codec.encode_batch_device(px, streams, off, flags, st)
        codec.decode_batch_device(streams, off, out, flags, st)
    codec.check(st)
    assert torch.equal(out, px), "round trip is not lossless"
    comp = int(off[-1].item())
    if kernel_times:
        codec.kernel_times()
        codec.set_kernel_timing(True)
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
    launches0 = codec.launches
    D.barrier()
    for a, m, b in ev:
        if flush is not None:
            flush.fill_(1)
        a.record()
        codec.encode_batch_device(px, streams, off, flags, st)
        m.record()
        codec.decode_batch_device(streams, off, out, flags, st)
        b.record()
    D.barrier()
    launches = codec.launches - launches0
    kt = None
    if kernel_times:
        codec.set_kernel_timing(False)
        kt = codec.kernel_times()
    codec.check(st)
    enc = sum(a.elapsed_time(m) for a, m, _ in ev) / steps
    dec = sum(m.elapsed_time(b) for _, m, b in ev) / steps
    tot, enc, dec = D.max([enc + dec, enc, dec])
    del streams, out, flush
    return {"ms": tot, "enc_ms": enc, "dec_ms": dec, "comp": comp, "raw": raw, "kernels": kt, "launches": launches,
            "l2": "inputs_exceed_l2" if raw >= (512 << 20) else "l2_flushed_between_steps"}


def secondary_lines(D, codec, args, hbm):
    """The other BASELINE.json configs, device-resident, few steps each (the whole default run stays within minutes)."""
    import flic_b200
    torch = D.torch
    sh, wl = flic_b200.sharding, flic_b200.workloads
    K, W = max(3, min(args.steps, 5)), 3
    out = {}

    def line(res, total_raw, extra):
        d = {"value_GBps": round(total_raw / (res["ms"] * 1e-3) / 1e9, 2),
             "encode_GBps": round(total_raw / (res["enc_ms"] * 1e-3) / 1e9, 2),
             "decode_GBps": round(total_raw / (res["dec_ms"] * 1e-3) / 1e9, 2), "ms_per_step": round(res["ms"], 4),
             "steps": K, "l2": res["l2"]}
        d.update(extra)
        return d

    def fracs(res):
        alg = (res["raw"] + res["comp"]) / 1e9
        return {"compressed_ratio": round(res["comp"] / res["raw"], 4),
                "encode_path_frac": round(alg / (res["enc_ms"] * 1e-3) / hbm, 4),
                "decode_path_frac": round(alg / (res["dec_ms"] * 1e-3) / hbm, 4)}

    # C3: the FIXED batch of 1024 x 1080p RGB split over the ranks (strong scaling)
    lo, hi = sh.batch_slice(1024, D.rank, D.world)
    uniq = torch.from_numpy(wl.make_batch("C3", n=DISTINCT)).cuda()
    px = uniq[torch.arange(lo, hi, device="cuda") % DISTINCT]
    res = device_roundtrip(D, codec, px, args.flags, K, W)
    out["C3_1024x1080p_rgb"] = line(res, 1024 * uniq[0].numel(), {"scaling": "strong", "images_per_rank": hi - lo,
                                                                   "distinct_images": DISTINCT, **fracs(res)})
    del px, uniq
    # C5: 64 x 4K RGBA uniform noise per rank (weak), the worst case for code lengths and for decode
    uniq = torch.from_numpy(wl.make_batch("C5", n=DISTINCT)).cuda()
    px = uniq[torch.arange(64, device="cuda") % DISTINCT]
    res = device_roundtrip(D, codec, px, args.flags, K, W)
    out["C5_64x4k_rgba_noise"] = line(res, D.world * px.numel(), {"scaling": "weak", "images_per_rank": 64,
                                                                  "distinct_images": DISTINCT, **fracs(res)})
    del px, uniq
    # the two pessimistic layouts on the main workload
    uniq = torch.from_numpy(wl.make_batch("C2", n=DISTINCT)).cuda()
    px = uniq[torch.arange(64, device="cuda") % DISTINCT]
    for name, fl in (("C2x64_one_stream_per_block", 0x20), ("C2x64_exact_sizes_lookback", 0x40)):
        res = device_roundtrip(D, codec, px, (args.flags & 0x1F) | fl, K, W)
        out[name] = line(res, D.world * px.numel(), {"scaling": "weak", "flags": (args.flags & 0x1F) | fl, **fracs(res)})
    del px, uniq
    torch.cuda.empty_cache()
    out["C4_16384sq_rgba_split"] = c4_split(D, codec, args, K, W)
    return out


def synth_rows_cuda(torch, w, h, c, seed, y0, y1):
    """Rows [y0, y1) of a w x h gradient+noise RGBA image, synthesised on the GPU (a 1 GiB image takes the host
    half a minute): same recipe as workloads.gradient_noise (diagonal ramp + N(0, 4) noise, opaque alpha), torch's
    generator instead of numpy's, seeded per 32-row band so any rank can make exactly its rows."""
    rows = torch.empty((y1 - y0, w, c), dtype=torch.uint8, device="cuda")
    x = torch.arange(w, dtype=torch.float32, device="cuda")[None, :, None]
    g = torch.Generator(device="cuda")
    for b0 in range(y0 - y0 % 32, y1, 32):
        lo, hi = max(b0, y0), min(b0 + 32, y1, h)
        if hi <= lo:
            continue
        g.manual_seed(seed * 1000003 + b0 // 32)
        noise = torch.randn((32, w, min(c, 3)), generator=g, device="cuda") * 4.0
        y = torch.arange(lo, hi, dtype=torch.float32, device="cuda")[:, None, None]
        ramp = 255.0 * (x + y) / max(w + h - 2, 1)
        base = torch.cat([ramp, 255.0 - ramp, ramp], dim=2)[..., : min(c, 3)]
        rows[lo - y0: hi - y0, :, : min(c, 3)] = (base + noise[lo - b0: hi - b0]).clamp_(0, 255).to(torch.uint8)
        if c == 4:
            rows[lo - y0: hi - y0, :, 3] = 255
    return rows


def c4_split(D, codec, args, K, W):
    """Config C4: ONE 16384x16384 RGBA8 image split by block rows over the ranks (strong scaling).  Timed region of a
    step: every rank encodes its rows; all-gather of (n_blocks, payload_words); payloads and directory entries travel
    by NCCL send/recv over NVLink straight into the spliced stream on rank 0; one kernel writes the header and
    rebases the directory; then the inverse: rank 0 cuts the stream at block-row boundaries, sends the parts, every
    rank finishes its part stream and decodes its rows.  Verified: spliced stream == the stream one GPU makes of
    rows it is given (same bytes per block, checked through decode), decoded rows == input rows on every rank."""
    import flic_b200
    torch = D.torch
    _, w, h, c, _, seed = flic_b200.workloads.CONFIGS["C4"]
    if args.c4_height:
        h = args.c4_height
    # peer memory when there is more than one rank and the box offers it (torch symmetric memory over NVLink): the pack
    # kernels write straight into rank 0's buffer, nothing synchronises with the host.  Otherwise NCCL send/recv of parts.
    sc, how, why = None, "nccl", None
    if D.world > 1 and args.c4_mode != "nccl":
        try:
            sc = flic_b200.sharding.PeerImageCodec(codec, w, h, c, args.flags, D.dist, D.rank, D.world)
            how = "peer"
        except Exception as e:  # no symmetric memory on this box / torch build
            why = "%s: %s" % (type(e).__name__, str(e)[:120])
        agree = D.sum([1.0 if how == "peer" else 0.0])[0]
        if agree != D.world:   # all ranks or none
            sc, how = None, "nccl"
    if sc is None:
        sc = flic_b200.sharding.ShardedImageCodec(codec, w, h, c, args.flags, D.dist, D.rank, D.world)
    rows = synth_rows_cuda(torch, w, h, c, seed, sc.y0, sc.y1)[None]
    st = torch.cuda.current_stream().cuda_stream
    full = None
    for _ in range(W):
        full = sc.encode(rows, st)
        got = sc.decode(full, st)
    codec.check(st)
    ok = bool(torch.equal(got, rows))
    if how == "peer":
        comp = sc.stream_bytes() if D.rank == 0 else 0
    else:
        comp = int(full.numel()) if D.rank == 0 else 0
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if rows.numel() < (512 << 20) else None
    launches0 = codec.launches
    D.barrier()
    for a, m, b in ev:
        if flush is not None:
            flush.fill_(1)  # a rank's share of the image fits in L2 at 8 GPUs: flush between steps
        a.record()
        full = sc.encode(rows, st)
        m.record()
        sc.decode(full, st)
        b.record()
    D.barrier()
    launches = codec.launches - launches0
    codec.check(st)
    enc = sum(a.elapsed_time(m) for a, m, _ in ev) / K
    dec = sum(m.elapsed_time(b) for _, m, b in ev) / K
    tot, enc, dec = D.max([enc + dec, enc, dec])
    allok = D.sum([0.0 if ok else 1.0])[0] == 0.0
    comp = int(D.sum([float(comp)])[0])
    raw = w * h * c
    l2 = "l2_flushed_between_steps" if flush is not None else "inputs_exceed_l2"
    del sc, rows, flush
    torch.cuda.empty_cache()
    return {"l2": l2, "value_GBps": round(raw / (tot * 1e-3) / 1e9, 2), "encode_GBps": round(raw / (enc * 1e-3) / 1e9, 2),
            "decode_GBps": round(raw / (dec * 1e-3) / 1e9, 2), "ms_per_step": round(tot, 4), "steps": K, "scaling": "strong",
            "width": w, "height": h, "raw_bytes": raw, "compressed_ratio": round(comp / raw, 4),
            "split": ("none (one GPU)" if D.world == 1 else
                      ("block rows over %d ranks, PEER MEMORY; timed: plan (histograms, tables, slots) + device-side all-gather of "
                       "the payload sizes + pack kernels writing straight into rank 0's buffer over NVLink + header, then every "
                       "rank pulls its part out of rank 0's buffer, finishes and decodes it; no host synchronisation" % D.world)
                      if how == "peer" else
                      ("block rows over %d ranks; timed: encode + all-gather of (n_blocks, payload_words) + NCCL send/recv of "
                       "parts into the spliced stream + splice kernel, then cut + send/recv + finish + decode" % D.world)),
            "transport": how if D.world > 1 else None, "peer_memory_unavailable": why,
            "round_trip_verified_on_every_rank": bool(allok), "gpu_launches_this_rank": launches, "data": "synthetic (GPU generator)"}


def e2e_pipelined(D, codec, host_px, cap, n, flags, steps):
    """End to end through the host-buffer ABI on pinned buffers, software-pipelined across steps: while step k's
    streams are being decoded (H2D of streams, D2H of pixels), step k+1 is being encoded (H2D of pixels, D2H of
    streams).  Two stream buffers alternate.  Every step copies its pixels from host memory and reads its decoded
    pixels back into host memory; the timed region covers `steps` full encode+decode round trips."""
    import numpy as np
    import flic_b200
    torch = D.torch
    h_in = host_px.numpy()
    h_str = [torch.empty(cap, dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    h_off = [np.zeros(n + 1, dtype=np.uint64) for _ in range(2)]
    h_out = torch.empty_like(host_px).pin_memory().numpy()

    def run(k):
        # prologue: encode step 0; steady state: encode(i+1) || decode(i); epilogue: decode the last
        codec.encode_submit(h_in, flags, out=h_str[0], offsets=h_off[0]); codec.wait(flic_b200.OP_ENCODE)
        nbytes = 0
        for i in range(k):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < k:
                codec.encode_submit(h_in, flags, out=h_str[nxt], offsets=h_off[nxt])
            nbytes = int(h_off[cur][n])
            codec.decode_submit(h_str[cur][:nbytes], h_off[cur], h_out)
            codec.wait(flic_b200.OP_DECODE)
            if i + 1 < k:
                codec.wait(flic_b200.OP_ENCODE)
        return nbytes

    run(2)
    assert np.array_equal(h_out, h_in), "host round trip is not lossless"
    k = max(3, min(steps, 10))
    D.barrier()
    t0 = time.perf_counter()
    nbytes = run(k)
    D.torch.cuda.synchronize()
    dt = D.max([time.perf_counter() - t0])[0]
    raw = h_in.nbytes
    per_dir = (raw + nbytes) * k / dt / 1e9
    return {"value": D.world * raw * k / dt / 1e9, "unit": "GB/s", "steps": k,
            "h2d_bytes_per_step": int(raw + nbytes + 8 * (n + 1)), "d2h_bytes_per_step": int(nbytes + raw + 8 * (n + 1) + 8),
            "pcie_GBps_per_direction_per_gpu": round(per_dir, 2),
            "note": "flic_encode_submit + flic_decode_submit + flic_wait on pinned host buffers, encode of step k+1 overlapping "
                    "decode of step k; host wall clock, max over ranks"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation is behind the licensing gate and there is no Rust
    toolchain here, so this arm is unavailable.  What CAN be timed on the host — the scalar C model of the provisional
    FLP0 format — is reported under `cpu_model`, explicitly not as the reference."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    line = {"impl": "reference", "unavailable": GATE, "metric": METRIC, "value": None, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "config": {"workload": args.workload}}
    if not args.no_cpu:
        import flic_b200
        cfg, n = WORKLOADS[args.workload]
        batch = flic_b200.workloads.make_batch(cfg, n=min(n, DISTINCT))
        line["cpu_model"] = cpu_model_roundtrip(batch, seconds=15.0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2x64", choices=sorted(WORKLOADS))
    ap.add_argument("--flags", type=lambda s: int(s, 0), default=0x01)
    ap.add_argument("--encoder", default="auto", choices=["auto", "fused", "staged"])
    ap.add_argument("--c4-mode", default="auto", choices=["auto", "peer", "nccl"], help="transport of the C4 block-row split")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--c4-height", type=int, default=0, help="override C4's 16384 rows (smoke runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import flic_b200

    D = Dist()
    numa, all_cpus = bind_to_gpu_numa(D.local)
    flic_b200.build_library()
    codec = flic_b200.Codec(D.local)
    codec.set_encoder(args.encoder)

    if args.workload == "C4":
        sampler = ClockSampler(D.local) if D.rank == 0 else None
        res = c4_split(D, codec, args, args.steps, args.warmup)
        clocks = sampler.stop() if sampler else None
        if D.rank == 0:
            print(json.dumps({"metric": METRIC, "value": res["value_GBps"], "unit": "GB/s", "n_gpus": D.world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                              "config": {"workload": "C4", "format": FORMAT, "l2": "inputs_exceed_l2"}, "c4": res,
                              "gpu_launches": res["gpu_launches_this_rank"], "clocks": clocks,
                              "parity": "engine vs FLP0 CPU model only; vs reference: unpinned — licensing gate"}))
        D.close()
        return

    cfg, n = WORKLOADS[args.workload]
    # weak scaling: every rank owns its own `n` images (independent units, no data-path collective)
    batch = flic_b200.workloads.make_batch(cfg, n=n, distinct=DISTINCT)
    _, h, w, c = batch.shape
    raw = batch.nbytes
    host_px = torch.from_numpy(batch).pin_memory()
    px = host_px.cuda(non_blocking=True)

    sampler = ClockSampler(D.local) if D.rank == 0 else None
    main_res = device_roundtrip(D, codec, px, args.flags, args.steps, args.warmup, kernel_times=True)
    comp, r = main_res["comp"], main_res["comp"] / raw
    ms_per_step = main_res["ms"]
    value = D.world * raw / (ms_per_step * 1e-3) / 1e9

    e2e = None
    if not args.no_e2e:
        cap = n * flic_b200.max_stream_bytes(w, h, c)
        e2e = e2e_pipelined(D, codec, host_px, cap, n, args.flags, args.steps)
    del px
    torch.cuda.empty_cache()
    hbm, peak_src = peaks()
    secondary = None
    if not args.no_secondary and args.workload == "C2x64":
        try:
            secondary = secondary_lines(D, codec, args, hbm)
        except Exception as e:  # the headline line must survive a failure in a side workload
            secondary = {"error": f"{type(e).__name__}: {e}"[:300]}
    clocks = sampler.stop() if sampler else None  # sampled across warm-up, the device-timed steps, e2e and secondary
    if D.rank != 0:
        D.close()
        return

    ktimes = main_res["kernels"]
    alg = {"k_histograms": raw, "k_tables": 0, "k_slots": 0, "k_pack": raw + comp, "k_finalize": 0, "k_decode": raw + comp,
           "k_encode": raw + comp, "k_decode_one": raw + comp}
    kernels = {}
    step_ms_total = ms_per_step * args.steps
    for k, (ms, cnt) in ktimes.items():
        if cnt:
            avg = ms / cnt
            kernels[k] = {"avg_ms": round(avg, 4), "launches": cnt, "share_of_step": round(ms / step_ms_total, 4),
                          "alg_GBps": round(alg[k] / (avg * 1e-3) / 1e9, 1) if alg[k] else None,
                          "frac_of_hbm_peak": round(alg[k] / (avg * 1e-3) / 1e9 / hbm, 4) if alg[k] else None}
    enc_frac = (raw + comp) / (main_res["enc_ms"] * 1e-3) / 1e9 / hbm
    dec_frac = (raw + comp) / (main_res["dec_ms"] * 1e-3) / 1e9 / hbm
    # the dominant kernel: the longest kernel of the slower path
    enc_k = [k for k in kernels if k in ("k_encode", "k_histograms", "k_tables", "k_pack", "k_slots", "k_finalize")]
    dec_k = [k for k in kernels if k in ("k_decode", "k_decode_one")]
    slow = enc_k if enc_frac <= dec_frac else dec_k
    dom = max(slow, key=lambda k: kernels[k]["avg_ms"]) if slow else None
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02b_traffic_c2x64.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload:
            for k in kernels:
                if k in tj["kernels"]:
                    kernels[k]["dram_traffic_bytes"] = tj["kernels"][k]["traffic_bytes"]
            if dom in tj["kernels"]:
                traffic, traffic_src = tj["kernels"][dom]["traffic_bytes"], "profiles/r02b_traffic_c2x64.json (ncu dram__bytes_read+write per launch)"
    except Exception:
        pass
    path_frac = min(enc_frac, dec_frac)
    roof = {"bound": "hbm", "kernel": dom, "achieved": round(path_frac * hbm, 1), "peak": hbm, "unit": "GB/s",
            "frac": round(path_frac, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "definition": "frac = min(encode path, decode path) of (1 + r) * N algorithmic bytes per pass / path time / peak; "
                          "`kernel` is the longest kernel of the slower path (its own fraction is in kernels[...])",
            "algorithmic_bytes_per_pass": raw + comp,
            "encode_path_frac": round(enc_frac, 4), "decode_path_frac": round(dec_frac, 4)}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": D.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_gpu": n, "distinct_images": min(n, DISTINCT), "width": w, "height": h,
                   "channels": c, "raw_bytes_per_gpu_step": raw, "flags": args.flags, "encoder": args.encoder,
                   "l2": main_res["l2"], "format": FORMAT, "host_numa": numa},
        "encode_GBps": round(D.world * raw / (main_res["enc_ms"] * 1e-3) / 1e9, 2),
        "decode_GBps": round(D.world * raw / (main_res["dec_ms"] * 1e-3) / 1e9, 2),
        "compressed_ratio": round(r, 4), "bits_per_pixel": round(8 * comp / (n * w * h), 3),
        "parity": "byte-exact vs FLP0 CPU model (tests/); vs reference: unpinned — licensing gate",
        "roofline": roof, "kernels": kernels, "gpu_launches": main_res["launches"], "clocks": clocks, "e2e": e2e,
        "secondary": secondary,
        "cpu_baseline": {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": GATE},
    }
    if D.world == 1 and not args.no_cpu:
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)  # the CPU model may use every core again
        line["cpu_model"] = cpu_model_roundtrip(batch[:DISTINCT], seconds=20.0)
    print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    main()
