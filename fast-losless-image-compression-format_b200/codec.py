"""ctypes binding of include/flic_b200.h.

No CPU fallback: every codec call goes through libflicb200.so's CUDA kernels.
A missing library or a missing sm_100 device raises FlicError — it never
degrades to the CPU model under oracle/ (that is test infrastructure only).
"""
import ctypes as C
import os

import numpy as np

from .build import library_path

BLOCK_W, BLOCK_H, MAX_CODE_LEN = 128, 32, 10
PRED_LEFT, FLAG_SUBGREEN = 1, 0x10
FLAG_ONE_STREAM, FLAG_EXACT = 0x20, 0x40   # optional stream layouts (include/flic_b200.h)
OP_ENCODE, OP_DECODE = 0, 1
ENCODER_FUSED, ENCODER_STAGED, ENCODER_AUTO = 0, 1, 2

_ERR = {
    -1: "invalid argument", -2: "output buffer too small", -3: "malformed stream", -4: "CUDA error",
    -5: "no sm_100 CUDA device (there is no CPU fallback)", -6: "unsupported format feature",
    -7: "device-side consistency check failed", -8: "a submitted operation has not been waited for",
}


class FlicError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__(f"flic error {code}: {_ERR.get(code, 'unknown')}" + (f" ({detail})" if detail else ""))


class _Info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("width", "height", "channels", "flags", "block_w", "block_h", "n_blocks", "payload_words")]


_lib = None


def load_library():
    """dlopen the in-tree library; fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise FlicError(-5, f"{path} not built — run __graft_entry__.build(); there is no CPU fallback")
    lib = C.CDLL(path)
    u32, u64, vp, i32 = C.c_uint32, C.c_uint64, C.c_void_p, C.c_int
    sig = {
        "flic_create": (i32, [i32, C.POINTER(vp)]),
        "flic_destroy": (None, [vp]),
        "flic_strerror": (C.c_char_p, [i32]),
        "flic_last_error": (C.c_char_p, [vp]),
        "flic_version": (i32, []),
        "flic_blocks_per_image": (u64, [u32, u32]),
        "flic_max_stream_bytes": (u64, [u32, u32, u32]),
        "flic_encode_batch_device": (i32, [vp, vp, u32, u32, u32, u32, u32, vp, u64, vp, vp]),
        "flic_decode_batch_device": (i32, [vp, vp, vp, u32, u32, u32, u32, u32, vp, vp]),
        "flic_check": (i32, [vp, vp]),
        "flic_encode_batch": (i32, [vp, vp, u32, u32, u32, u32, u32, vp, u64, vp]),
        "flic_decode_batch": (i32, [vp, vp, vp, u32, vp, u64]),
        "flic_peek": (i32, [vp, u64, C.POINTER(_Info)]),
        "flic_splice_block_rows": (i32, [C.POINTER(vp), C.POINTER(u64), u32, vp, u64, C.POINTER(u64)]),
        "flic_stage_histograms": (i32, [vp, vp, u32, u32, u32, u32, u32, vp, vp, vp]),
        "flic_stage_tables": (i32, [vp, vp, u64, vp, vp, vp]),
        "flic_launch_count": (u64, [vp]),
        "flic_set_kernel_timing": (i32, [vp, i32]),
        "flic_get_kernel_times": (i32, [vp, C.POINTER(C.c_double), C.POINTER(u64)]),
        "flic_set_option": (i32, [vp, i32, i32]),
        "flic_get_phase_clocks": (i32, [vp, C.POINTER(u64)]),
        "flic_host_register": (i32, [vp, u64]),
        "flic_host_unregister": (i32, [vp]),
        "flic_encode_submit": (i32, [vp, vp, u32, u32, u32, u32, u32, vp, u64, vp]),
        "flic_decode_submit": (i32, [vp, vp, vp, u32, vp, u64]),
        "flic_wait": (i32, [vp, i32]),
        "flic_splice_block_rows_device": (i32, [vp, C.POINTER(vp), C.POINTER(u64), u32, vp, u64, C.POINTER(u64), vp]),
        "flic_splice_plan": (i32, [C.POINTER(u32), C.POINTER(u32), u32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "flic_encode_plan_device": (i32, [vp, vp, u32, u32, u32, u32, vp, vp]),
        "flic_encode_emit_device": (i32, [vp, vp, u64, u32, u32, vp, vp]),
        "flic_splice_header_device": (i32, [vp, vp, u64, u32, u32, u32, u32, vp, vp]),
        "flic_pull_part_device": (i32, [vp, vp, u64, u32, u32, u32, vp, u64, vp, vp]),
        "flic_splice_finish_device": (i32, [vp, vp, C.POINTER(u32), C.POINTER(u32), u32, u32, u32, u32, u32, vp]),
        "flic_split_finish_device": (i32, [vp, vp, u32, u32, u32, u32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


EXPORTED = (
    "flic_create flic_destroy flic_strerror flic_last_error flic_version flic_blocks_per_image "
    "flic_max_stream_bytes flic_encode_batch_device flic_decode_batch_device flic_check flic_encode_batch "
    "flic_decode_batch flic_peek flic_splice_block_rows flic_stage_histograms flic_stage_tables "
    "flic_launch_count flic_set_kernel_timing flic_get_kernel_times flic_set_option flic_get_phase_clocks flic_host_register "
    "flic_host_unregister flic_encode_submit flic_decode_submit flic_wait flic_splice_block_rows_device "
    "flic_splice_plan flic_splice_finish_device flic_split_finish_device flic_encode_plan_device "
    "flic_encode_emit_device flic_splice_header_device flic_pull_part_device"
).split()

KERNELS = ("k_histograms", "k_tables", "k_pack", "k_finalize", "k_decode", "k_slots", "k_encode", "k_decode_one")


def max_stream_bytes(w, h, c):
    return int(load_library().flic_max_stream_bytes(w, h, c))


def blocks_per_image(w, h):
    return ((w + BLOCK_W - 1) // BLOCK_W) * ((h + BLOCK_H - 1) // BLOCK_H)


def peek(stream):
    """Parse a stream header -> dict (host only, no GPU)."""
    buf = np.ascontiguousarray(np.frombuffer(stream, dtype=np.uint8))
    info = _Info()
    rc = load_library().flic_peek(buf.ctypes.data, buf.size, C.byref(info))
    if rc:
        raise FlicError(rc)
    return {n: getattr(info, n) for n, _ in _Info._fields_}


def splice_block_rows(parts):
    """Concatenate streams that each cover a run of whole block rows of one image (host only)."""
    lib = load_library()
    bufs = [np.ascontiguousarray(np.frombuffer(p, dtype=np.uint8)) for p in parts]
    k = len(bufs)
    ptrs = (C.c_void_p * k)(*[b.ctypes.data for b in bufs])
    sizes = (C.c_uint64 * k)(*[b.size for b in bufs])
    out = np.empty(sum(b.size for b in bufs), dtype=np.uint8)
    n = C.c_uint64(0)
    rc = lib.flic_splice_block_rows(ptrs, sizes, k, out.ctypes.data, out.size, C.byref(n))
    if rc:
        raise FlicError(rc)
    return out[: n.value].copy()


def splice_plan(part_blocks, part_payload_words):
    """Where each part's directory entries and payload land in the spliced stream (host arithmetic only).
    Returns (dir_byte_offsets, payload_byte_offsets, total_bytes)."""
    k = len(part_blocks)
    nb = (C.c_uint32 * k)(*[int(x) for x in part_blocks])
    pw = (C.c_uint32 * k)(*[int(x) for x in part_payload_words])
    d, p, tot = (C.c_uint64 * k)(), (C.c_uint64 * k)(), C.c_uint64(0)
    rc = load_library().flic_splice_plan(nb, pw, k, d, p, C.byref(tot))
    if rc:
        raise FlicError(rc)
    return list(d), list(p), int(tot.value)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class Codec:
    """One engine context on one GPU. Thread-compatible, not thread-safe (like the C ABI)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.flic_create(int(device), C.byref(h))
        if rc:
            raise FlicError(rc)
        self.h, self.device = h, int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.flic_destroy(self.h)
            self.h = None

    __del__ = close

    def _chk(self, rc):
        if rc:
            raise FlicError(rc, (self.lib.flic_last_error(self.h) or b"").decode())

    @property
    def launches(self):
        return int(self.lib.flic_launch_count(self.h))

    def set_encoder(self, which):
        """'fused' (one pass over the pixels), 'staged' (five-kernel pipeline) or 'auto' (default: by job size);
        all give the same bytes."""
        self._chk(self.lib.flic_set_option(self.h, 1, {"fused": ENCODER_FUSED, "staged": ENCODER_STAGED, "auto": ENCODER_AUTO}[which]))

    # ---- submit / wait: one encode and one decode may be in flight together ----
    def encode_submit(self, pixels, flags=PRED_LEFT, out=None, offsets=None):
        """Asynchronous encode_batch on caller-owned (ideally pinned) buffers; returns (out, offsets) to be read
        after wait(OP_ENCODE)."""
        px = pixels
        if px.dtype != np.uint8 or not px.flags.c_contiguous or px.ndim != 4:
            raise FlicError(-1, "pixels must be a C-contiguous uint8 [n,h,w,c] array that outlives the wait")
        n, h, w, c = px.shape
        if out is None:
            out = np.empty(max(n * max_stream_bytes(w, h, c), 1), dtype=np.uint8)
        if offsets is None:
            offsets = np.zeros(n + 1, dtype=np.uint64)
        self._keep_enc = (px, out, offsets)
        self._chk(self.lib.flic_encode_submit(self.h, px.ctypes.data, n, w, h, c, flags, out.ctypes.data, out.size,
                                              offsets.ctypes.data))
        return out, offsets

    def decode_submit(self, streams, offsets, out):
        if streams.dtype != np.uint8 or offsets.dtype != np.uint64 or not out.flags.c_contiguous:
            raise FlicError(-1, "streams uint8, offsets uint64, out C-contiguous; all must outlive the wait")
        self._keep_dec = (streams, offsets, out)
        self._chk(self.lib.flic_decode_submit(self.h, streams.ctypes.data, offsets.ctypes.data, offsets.size - 1,
                                              out.ctypes.data, out.nbytes))
        return out

    def wait(self, op):
        self._chk(self.lib.flic_wait(self.h, int(op)))

    # ---- host-buffer API: numpy in, numpy out; H2D/D2H inside the call ----
    def encode_batch(self, pixels, flags=PRED_LEFT, out=None, offsets=None):
        """pixels: uint8 [n,h,w,c] (C-contiguous). Returns (streams uint8[total], offsets uint64[n+1])."""
        px = np.ascontiguousarray(pixels, dtype=np.uint8)
        if px.ndim != 4:
            raise FlicError(-1, "pixels must be [n,h,w,c]")
        n, h, w, c = px.shape
        cap = n * max_stream_bytes(w, h, c) if (n and h and w and 1 <= c <= 4) else 0
        if out is None:
            out = np.empty(max(cap, 1), dtype=np.uint8)
        if offsets is None:
            offsets = np.zeros(n + 1, dtype=np.uint64)
        self._chk(self.lib.flic_encode_batch(self.h, px.ctypes.data, n, w, h, c, flags, out.ctypes.data, out.size,
                                             offsets.ctypes.data))
        return out[: int(offsets[n])], offsets

    def decode_batch(self, streams, offsets, out=None):
        """streams: uint8[total], offsets: uint64[n+1]. Returns uint8 [n,h,w,c]."""
        s = np.ascontiguousarray(streams, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = off.size - 1
        if n < 1:
            raise FlicError(-1, "empty batch")
        info = peek(s[int(off[0]): int(off[1])])
        if out is None:
            out = np.empty((n, info["height"], info["width"], info["channels"]), dtype=np.uint8)
        self._chk(self.lib.flic_decode_batch(self.h, s.ctypes.data, off.ctypes.data, n, out.ctypes.data, out.nbytes))
        return out

    def encode(self, image, flags=PRED_LEFT):
        s, _ = self.encode_batch(np.asarray(image)[None], flags)
        return s.copy()

    def decode(self, stream):
        s = np.ascontiguousarray(np.frombuffer(stream, dtype=np.uint8))
        return self.decode_batch(s, np.array([0, s.size], dtype=np.uint64))[0]

    # ---- device-resident API: torch CUDA tensors, asynchronous on `stream` ----
    def encode_batch_device(self, pixels, streams, offsets, flags=PRED_LEFT, stream=0):
        n, h, w, c = pixels.shape
        self._chk(self.lib.flic_encode_batch_device(self.h, _ptr(pixels), n, w, h, c, flags, _ptr(streams),
                                                    streams.numel(), _ptr(offsets), C.c_void_p(stream)))

    def decode_batch_device(self, streams, offsets, pixels, flags=PRED_LEFT, stream=0):
        n, h, w, c = pixels.shape
        self._chk(self.lib.flic_decode_batch_device(self.h, _ptr(streams), _ptr(offsets), n, w, h, c, flags,
                                                    _ptr(pixels), C.c_void_p(stream)))

    def check(self, stream=0):
        self._chk(self.lib.flic_check(self.h, C.c_void_p(stream)))

    # ---- block-row splice on the device (torch CUDA uint8 tensors) ----
    def splice_block_rows_device(self, parts, out, stream=0):
        """parts: list of CUDA uint8 tensors, each one stream of a run of whole block rows; out: CUDA uint8 tensor.
        Returns the number of bytes of the spliced stream written to out."""
        k = len(parts)
        ptrs = (C.c_void_p * k)(*[p.data_ptr() for p in parts])
        sizes = (C.c_uint64 * k)(*[p.numel() for p in parts])
        n = C.c_uint64(0)
        self._chk(self.lib.flic_splice_block_rows_device(self.h, ptrs, sizes, k, _ptr(out), out.numel(), C.byref(n),
                                                         C.c_void_p(stream)))
        return int(n.value)

    def splice_finish_device(self, out, part_blocks, part_payload_words, w, h_total, c, flags=PRED_LEFT, stream=0):
        k = len(part_blocks)
        nb = (C.c_uint32 * k)(*[int(x) for x in part_blocks])
        pw = (C.c_uint32 * k)(*[int(x) for x in part_payload_words])
        self._chk(self.lib.flic_splice_finish_device(self.h, _ptr(out), nb, pw, k, w, h_total, c, flags, C.c_void_p(stream)))

    # ---- block-row split with peer memory (stream buffers are raw device addresses: they may be another GPU's) ----
    def encode_plan_device(self, rows, flags, d_payload_words, stream=0):
        """rows: [1, h, w, c] CUDA uint8; d_payload_words: CUDA int64[>=1] that receives the part's payload words."""
        _, h, w, c = rows.shape
        self._chk(self.lib.flic_encode_plan_device(self.h, _ptr(rows), w, h, c, flags, _ptr(d_payload_words), C.c_void_p(stream)))

    def encode_emit_device(self, stream_ptr, capacity_bytes, total_blocks, first_block, d_base_words, stream=0):
        self._chk(self.lib.flic_encode_emit_device(self.h, C.c_void_p(int(stream_ptr)), capacity_bytes, total_blocks, first_block,
                                                   _ptr(d_base_words), C.c_void_p(stream)))

    def splice_header_device(self, stream_ptr, capacity_bytes, w, h_total, c, flags, d_total_words, stream=0):
        self._chk(self.lib.flic_splice_header_device(self.h, C.c_void_p(int(stream_ptr)), capacity_bytes, w, h_total, c, flags,
                                                     _ptr(d_total_words), C.c_void_p(stream)))

    def pull_part_device(self, stream_ptr, stream_bytes, total_blocks, first_block, part_blocks, part, d_part_bytes, stream=0):
        self._chk(self.lib.flic_pull_part_device(self.h, C.c_void_p(int(stream_ptr)), stream_bytes, total_blocks, first_block, part_blocks,
                                                 _ptr(part), part.numel(), _ptr(d_part_bytes), C.c_void_p(stream)))

    def split_finish_device(self, part, w, h_part, c, flags=PRED_LEFT, stream=0):
        self._chk(self.lib.flic_split_finish_device(self.h, _ptr(part), w, h_part, c, flags, C.c_void_p(stream)))

    # ---- measurement hooks ----
    def set_kernel_timing(self, enable):
        self._chk(self.lib.flic_set_kernel_timing(self.h, int(bool(enable))))

    def kernel_times(self):
        """{kernel: (total_ms, launches)} since the last call (waits for the recorded events)."""
        ms = (C.c_double * len(KERNELS))()
        cnt = (C.c_uint64 * len(KERNELS))()
        self._chk(self.lib.flic_get_kernel_times(self.h, ms, cnt))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNELS)}

    def phase_clocks(self):
        """Per-phase SM-cycle sums of k_encode [0:8] and k_decode_one [8:16] since the last call (needs
        FLIC_PHASE_CLOCKS=1 at creation)."""
        cyc = (C.c_uint64 * 16)()
        self._chk(self.lib.flic_get_phase_clocks(self.h, cyc))
        return [int(x) for x in cyc]

    # ---- stage-level (parity tests) ----
    def stage_histograms(self, pixels, hist, flags=PRED_LEFT, stream=0, flat=None):
        n, h, w, c = pixels.shape
        self._chk(self.lib.flic_stage_histograms(self.h, _ptr(pixels), n, w, h, c, flags, _ptr(hist),
                                                 _ptr(flat) if flat is not None else None, C.c_void_p(stream)))

    def stage_tables(self, hist, table, stream=0, bits=None):
        self._chk(self.lib.flic_stage_tables(self.h, _ptr(hist), hist.shape[0], _ptr(table),
                                             _ptr(bits) if bits is not None else None, C.c_void_p(stream)))
