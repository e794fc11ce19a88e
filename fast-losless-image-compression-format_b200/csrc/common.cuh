// common.cuh — geometry, format constants and the in-register row-residual
// routine shared by the histogram and pack kernels.  sm_100a only.
//
// Nothing here follows the reference's source (licensing gate, LICENSING.md);
// the format is the provisional FLP0 bitstream specified in DESIGN.md.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flic_b200.h"

namespace flic {

constexpr int kBW = FLIC_BLOCK_W;            // pixels per block row
constexpr int kBH = FLIC_BLOCK_H;            // rows per block == lanes per decode warp
constexpr int kL = FLIC_MAX_CODE_LEN;        // max code length
constexpr int kLutSize = 1 << kL;
constexpr int kHdrWords = 8;                 // 32-byte stream header
constexpr int kBlkHdrWords = 32 + kBH / 2;   // 256 length nibbles + 32 u16 row word counts
constexpr int kRowWordsMax = (kBW * 4 * kL + 31) / 32;  // 176: worst-case words of one row sub-stream
constexpr uint32_t kMagic = 0x30504C46u;
constexpr uint32_t kLenSole = 15;

// device-side error bits, OR-ed into ctx->d_err[0]
constexpr uint32_t kErrCapacity = 1u;
constexpr uint32_t kErrWatchdog = 2u;
constexpr uint32_t kErrFormat = 4u;

struct Geo {
    uint32_t n, w, h, c, flags;
    uint32_t nbx, nby, nb;       // blocks per image (x, y, total)
    uint64_t pitch, img_stride;  // bytes
    uint32_t aligned16;          // base, pitch and image stride are all 16-byte multiples
};

struct BlockPos {
    uint32_t img, b, x0, y0, bwa, bha, rb;  // rb = bytes per block row = bwa * c
};

__device__ __forceinline__ BlockPos block_pos(const Geo &g, uint64_t gb) {
    BlockPos p;
    p.img = (uint32_t)(gb / g.nb);
    p.b = (uint32_t)(gb - (uint64_t)p.img * g.nb);
    uint32_t by = p.b / g.nbx, bx = p.b - by * g.nbx;
    p.x0 = bx * kBW;
    p.y0 = by * kBH;
    p.bwa = min((uint32_t)kBW, g.w - p.x0);
    p.bha = min((uint32_t)kBH, g.h - p.y0);
    p.rb = p.bwa * g.c;
    return p;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 16 bytes of a block row starting at byte `off`; bytes at or past `rb` read as 0.
__device__ __forceinline__ uint4 load_chunk16(const uint8_t *row, int off, int rb, bool aligned) {
    int nv = rb - off;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (nv <= 0) return v;
    if (aligned && nv >= 16) return ldg_nc_v4(row + off);
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (j < nv) w[j >> 2] |= (uint32_t)__ldg(row + off + j) << (8 * (j & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Packed byte-wise subtract-green for 3-channel data whose first byte has
// channel phase `p` (0,1,2): ch0 bytes take the next byte (G), ch2 the previous.
__device__ __forceinline__ uint32_t subgreen3(uint32_t prev, uint32_t cur, uint32_t next, int p) {
    uint32_t nb = __funnelshift_r(cur, next, 8);   // bytes i+1
    uint32_t pb = __funnelshift_r(prev, cur, 24);  // bytes i-1
    uint32_t m0 = p == 0 ? 0xFF0000FFu : (p == 1 ? 0x00FF0000u : 0x0000FF00u);
    uint32_t m2 = p == 0 ? 0x00FF0000u : (p == 1 ? 0x0000FF00u : 0xFF0000FFu);
    return __vsub4(cur, (nb & m0) | (pb & m2));
}

__device__ __forceinline__ uint32_t subgreen4(uint32_t px) {
    uint32_t g = (px >> 8) & 0xFFu;
    return __vsub4(px, g | (g << 16));
}
__device__ __forceinline__ uint32_t addgreen4(uint32_t px) {
    uint32_t g = (px >> 8) & 0xFFu;
    return __vadd4(px, g | (g << 16));
}

// Residual bytes [16*lane, 16*lane+16) of block row `r` (one warp spans the
// row).  FLP0 §2: optional subtract-green, then pred = left pixel; at x == 0
// the pixel above; at (0,0) zero — all inside the block.  Must be called by
// all 32 lanes.  Returns the packed residuals; *nv = how many are real.
__device__ __forceinline__ uint4 row_residuals(const uint8_t *pixels, const Geo &g, const BlockPos &p,
                                               int r, int lane, int *nv) {
    const uint8_t *row = pixels + (uint64_t)p.img * g.img_stride + (uint64_t)(p.y0 + r) * g.pitch +
                         (uint64_t)p.x0 * g.c;
    const int c = (int)g.c;
    const bool sg = (g.flags & FLIC_FLAG_SUBGREEN) && c >= 3;
    uint4 v = load_chunk16(row, 16 * lane, (int)p.rb, g.aligned16 != 0);
    *nv = max(0, min(16, (int)p.rb - 16 * lane));

    // transformed pixel above (lane 0 only), packed in the low c bytes
    uint32_t up = 0;
    if (lane == 0 && r > 0) {
        const uint8_t *u = row - g.pitch;
        uint32_t b0 = __ldg(u), b1 = c > 1 ? __ldg(u + 1) : 0u, b2 = c > 2 ? __ldg(u + 2) : 0u,
                 b3 = c > 3 ? __ldg(u + 3) : 0u;
        if (sg) { b0 = (b0 - b1) & 0xFFu; b2 = (b2 - b1) & 0xFFu; }
        up = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }

    if (sg) {
        if (c == 4) {
            v.x = subgreen4(v.x); v.y = subgreen4(v.y); v.z = subgreen4(v.z); v.w = subgreen4(v.w);
        } else {
            uint32_t pw = __shfl_up_sync(0xFFFFFFFFu, v.w, 1);
            uint32_t nw = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
            int ph = lane % 3;  // (16*lane) % 3
            uint4 t;
            t.x = subgreen3(pw, v.x, v.y, ph);
            t.y = subgreen3(v.x, v.y, v.z, (ph + 1) % 3);
            t.z = subgreen3(v.y, v.z, v.w, (ph + 2) % 3);
            t.w = subgreen3(v.z, v.w, nw, ph);
            v = t;
        }
    }

    const int sh = 8 * (4 - c);
    uint32_t pl = __shfl_up_sync(0xFFFFFFFFu, v.w, 1);
    if (lane == 0) pl = up << sh;
    uint4 res;
    res.x = __vsub4(v.x, __funnelshift_r(pl, v.x, sh));
    res.y = __vsub4(v.y, __funnelshift_r(v.x, v.y, sh));
    res.z = __vsub4(v.z, __funnelshift_r(v.y, v.z, sh));
    res.w = __vsub4(v.w, __funnelshift_r(v.z, v.w, sh));
    return res;
}

// Fast path of row_residuals for a full-width (128-pixel) block row of a 16-byte-aligned image:
// channels and the colour transform are compile-time, `row` already points at the lane's chunk.
template <int C, bool SG>
__device__ __forceinline__ uint4 row_residuals_fast(const uint8_t *row, uint64_t pitch, int r, int lane, int *nv) {
    constexpr int kLanes = kBW * C / 16;  // lanes that own bytes of the row
    const bool own = lane < kLanes;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (own) v = ldg_nc_v4(row);
    *nv = own ? 16 : 0;
    uint32_t up = 0;
    if (lane == 0 && r > 0) {
        up = __ldg(reinterpret_cast<const uint32_t *>(row - pitch));  // 4-byte aligned: row is 16-byte aligned
        if (C < 4) up &= (1u << (8 * (C & 3))) - 1u;
        if (SG && C >= 3) up = subgreen4(up);  // bytes 0,2 minus byte 1; byte 3 (alpha or masked) untouched
    }
    if (SG && C == 4) {
        v.x = subgreen4(v.x); v.y = subgreen4(v.y); v.z = subgreen4(v.z); v.w = subgreen4(v.w);
    } else if (SG && C == 3) {
        const uint32_t pw = __shfl_up_sync(0xFFFFFFFFu, v.w, 1), nw = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
        const int ph = lane % 3;
        uint4 t;
        t.x = subgreen3(pw, v.x, v.y, ph);
        t.y = subgreen3(v.x, v.y, v.z, (ph + 1) % 3);
        t.z = subgreen3(v.y, v.z, v.w, (ph + 2) % 3);
        t.w = subgreen3(v.z, v.w, nw, ph);
        v = t;
    }
    constexpr int sh = 8 * (4 - C);
    uint32_t pl = __shfl_up_sync(0xFFFFFFFFu, v.w, 1);
    if (lane == 0) pl = up << sh;
    uint4 res;
    res.x = __vsub4(v.x, __funnelshift_r(pl, v.x, sh));
    res.y = __vsub4(v.y, __funnelshift_r(v.x, v.y, sh));
    res.z = __vsub4(v.z, __funnelshift_r(v.y, v.z, sh));
    res.w = __vsub4(v.w, __funnelshift_r(v.z, v.w, sh));
    return res;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 16);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 8);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
    return v;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// launchers (defined in the .cu files, used by api.cu)
void launch_histograms(const uint8_t *d_pixels, const Geo &g, uint16_t *d_hist, uint4 *d_resid, cudaStream_t s);
void launch_tables(const uint16_t *d_hist, uint64_t nblocks, uint16_t *d_table, cudaStream_t s);
void launch_pack(const uint4 *d_resid, const Geo &g, const uint16_t *d_table, uint32_t *d_streams,
                 uint64_t capacity_words, unsigned long long *d_status, unsigned long long *d_dirE,
                 uint32_t *d_err, cudaStream_t s);
void launch_finalize(const Geo &g, const unsigned long long *d_dirE, uint32_t *d_streams,
                     uint64_t capacity_words, unsigned long long *d_offsets, uint32_t *d_err,
                     cudaStream_t s);
void launch_decode(const uint32_t *d_streams, const unsigned long long *d_offsets, const Geo &g,
                   uint8_t *d_pixels, uint32_t *d_err, cudaStream_t s);

}  // namespace flic
