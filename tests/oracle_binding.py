"""ctypes binding of oracle/libflp0_oracle.so — test infrastructure only."""
import ctypes as C

import numpy as np


class Oracle:
    BW, BH = 128, 32

    def __init__(self, path):
        lib = C.CDLL(path)
        u32, vp = C.c_uint32, C.c_void_p
        lib.flp0_max_stream_bytes.restype = C.c_size_t
        lib.flp0_max_stream_bytes.argtypes = [u32] * 5
        lib.flp0_encode.restype = C.c_int64
        lib.flp0_encode.argtypes = [vp, u32, u32, u32, u32, u32, u32, vp, C.c_size_t]
        lib.flp0_decode.restype = C.c_int
        lib.flp0_decode.argtypes = [vp, C.c_size_t, vp, C.c_size_t]
        lib.flp0_block_residuals.restype = C.c_size_t
        lib.flp0_block_residuals.argtypes = [vp] + [u32] * 8 + [vp]
        lib.flp0_block_histogram.restype = C.c_uint32
        lib.flp0_block_histogram.argtypes = [vp] + [u32] * 8 + [vp, vp]
        lib.flp0_build_lengths.restype = None
        lib.flp0_build_lengths.argtypes = [vp, vp]
        lib.flp0_assign_codes.restype = None
        lib.flp0_assign_codes.argtypes = [vp, vp]
        self.lib = lib

    def encode(self, img, flags=1, bw=BW, bh=BH):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        cap = self.lib.flp0_max_stream_bytes(w, h, c, bw, bh)
        out = np.empty(cap, dtype=np.uint8)
        n = self.lib.flp0_encode(img.ctypes.data, w, h, c, flags, bw, bh, out.ctypes.data, cap)
        if n < 0:
            raise ValueError(f"flp0_encode -> {n}")
        return out[:n].copy()

    def encode_rc(self, img, flags=1, bw=BW, bh=BH, cap=None):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        cap = self.lib.flp0_max_stream_bytes(w, h, c, bw, bh) if cap is None else cap
        out = np.empty(max(cap, 1), dtype=np.uint8)
        return int(self.lib.flp0_encode(img.ctypes.data, w, h, c, flags, bw, bh, out.ctypes.data, cap))

    def decode(self, stream, shape):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.zeros(shape, dtype=np.uint8)
        rc = self.lib.flp0_decode(s.ctypes.data, s.size, out.ctypes.data, out.size)
        if rc:
            raise ValueError(f"flp0_decode -> {rc}")
        return out

    def decode_rc(self, stream, shape):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = np.zeros(shape, dtype=np.uint8)
        return int(self.lib.flp0_decode(s.ctypes.data, s.size, out.ctypes.data, out.size))

    def block_histograms(self, img, flags=1, with_flat=False):
        """uint16 [n_blocks,256] residual histograms in block raster order, flat channels excluded;
        with_flat=True also returns uint32 [n_blocks,2] = (flat mask, packed flat values)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        nbx, nby = -(-w // self.BW), -(-h // self.BH)
        hist = np.zeros(256, dtype=np.uint32)
        val = np.zeros(4, dtype=np.uint8)
        out = np.zeros((nbx * nby, 256), dtype=np.uint16)
        meta = np.zeros((nbx * nby, 2), dtype=np.uint32)
        for by in range(nby):
            for bx in range(nbx):
                m = self.lib.flp0_block_histogram(img.ctypes.data, w, h, c, flags, bx * self.BW, by * self.BH,
                                                  self.BW, self.BH, hist.ctypes.data, val.ctypes.data)
                out[by * nbx + bx] = hist
                meta[by * nbx + bx] = (m, int(val.view(np.uint32)[0]))
        return (out, meta) if with_flat else out

    def table(self, hist):
        """hist uint[256] -> uint16[256] entries len<<12|code (15<<12 for a sole symbol), as k_tables emits."""
        hh = np.ascontiguousarray(hist, dtype=np.uint32)
        ln = np.zeros(256, dtype=np.uint8)
        cd = np.zeros(256, dtype=np.uint16)
        self.lib.flp0_build_lengths(hh.ctypes.data, ln.ctypes.data)
        self.lib.flp0_assign_codes(ln.ctypes.data, cd.ctypes.data)
        return (ln.astype(np.uint16) << 12) | cd

    def lengths(self, hist):
        hh = np.ascontiguousarray(hist, dtype=np.uint32)
        ln = np.zeros(256, dtype=np.uint8)
        self.lib.flp0_build_lengths(hh.ctypes.data, ln.ctypes.data)
        return ln
