// common.cuh — geometry, format constants, the in-register row-residual routine
// shared by the histogram kernels and the fused encoder, and the launchers the C
// ABI (api.cu) calls.  sm_100a only.
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
constexpr int kFlatWord = 32 + kBH / 2;     // block header word holding the flat-channel mask; the values follow
constexpr int kBlkHdrWords = kFlatWord + 2;  // 256 length nibbles + 32 u16 row word counts + flat mask + flat values
constexpr int kBlkHdrWords1 = 32 + 2;        // FLIC_FLAG_ONE_STREAM: 256 length nibbles + flat mask + flat values, then one bit stream
constexpr uint32_t kVersion = 3;
constexpr int kRowWordsMax = (kBW * 4 * kL + 31) / 32;  // 160: worst-case words of one row sub-stream
constexpr uint32_t kMagic = 0x30504C46u;
constexpr uint32_t kLenSole = 15;

// device-side error bits, OR-ed into ctx->d_err[0]
constexpr uint32_t kErrCapacity = 1u;
constexpr uint32_t kErrFormat = 4u;
constexpr uint32_t kErrSlot = 8u;  // a block outgrew the slot k_slots computed for it (internal error)
constexpr uint32_t kErrRange = 16u;  // an image's payload does not fit the container's u32 word offsets
constexpr uint32_t kErrLayout = 32u;  // a kernel's shared-memory layout assumption does not hold (internal error)

__host__ __device__ __forceinline__ bool one_stream(uint32_t flags) { return (flags & FLIC_FLAG_ONE_STREAM) != 0; }
__host__ __device__ __forceinline__ int blk_hdr_words(uint32_t flags) { return one_stream(flags) ? kBlkHdrWords1 : kBlkHdrWords; }

struct Geo {
    uint32_t n, w, h, c, flags;
    uint32_t nbx, nby, nb;       // blocks per image (x, y, total)
    uint64_t pitch, img_stride;  // bytes
    uint32_t aligned16;          // base, pitch and image stride are all 16-byte multiples
    uint32_t aligned32;          // ... and 32-byte multiples (256-bit stores in k_decode)
};

struct BlockPos {
    uint32_t img, b, x0, y0, bwa, bha, rb;  // rb = bytes per block row = bwa * c
};

__device__ __forceinline__ BlockPos block_pos(const Geo &g, uint64_t gb64) {
    BlockPos p;
    const uint32_t gb = (uint32_t)gb64;  // make_geo() keeps n * nb below 2^31: 32-bit divisions
    p.img = gb / g.nb;
    p.b = gb - p.img * g.nb;
    uint32_t by = p.b / g.nbx, bx = p.b - by * g.nbx;
    p.x0 = bx * kBW;
    p.y0 = by * kBH;
    p.bwa = min((uint32_t)kBW, g.w - p.x0);
    p.bha = min((uint32_t)kBH, g.h - p.y0);
    p.rb = p.bwa * g.c;
    return p;
}

// One CTA per block, two grid shapes.  (nbx, nby, n): the block's coordinates ARE the CTA's — no divisions (the two of
// block_pos are ~45 instructions per warp, 4-7 % of the encode kernels' instruction count).  n * nb x 1 x 1: the general
// case (more than 65535 block rows or images).  Both enumerate the blocks in the same order.
__host__ __device__ __forceinline__ bool grid3_ok(const Geo &g) { return g.nby <= 65535u && g.n <= 65535u; }
__device__ __forceinline__ BlockPos block_pos_cta(const Geo &g, bool grid3, uint64_t &gb) {
    if (!grid3) {
        gb = blockIdx.x;
        return block_pos(g, gb);
    }
    BlockPos p;
    p.img = blockIdx.z;
    p.b = blockIdx.y * g.nbx + blockIdx.x;
    gb = (uint64_t)p.img * g.nb + p.b;
    p.x0 = blockIdx.x * kBW;
    p.y0 = blockIdx.y * kBH;
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

// ---- encode-side lane mapping -------------------------------------------------------------
// One warp spans a block row; every lane owns exactly FOUR PIXELS of it, i.e. 4*C bytes = C
// 32-bit words, for any channel count (so a 384-byte RGB row keeps all 32 lanes busy, and four
// whole pixels per lane mean the colour transform never crosses a lane).

// The lane's C words of block row `row` (pointer to the block row's first byte); pixels at or
// past `bwa` read as 0.  *nv = how many of the lane's 4*C bytes are real.
template <int C>
__device__ __forceinline__ void load_lane_pixels(const uint8_t *row, int lane, int bwa, bool fast, uint32_t (&v)[C],
                                                 int *nv) {
    const int npx = max(0, min(4, bwa - 4 * lane));
    *nv = npx * C;
#pragma unroll
    for (int j = 0; j < C; ++j) v[j] = 0;
    if (fast && npx == 4) {  // 16-byte-aligned image: 4*C*lane is a multiple of C words
        const uint8_t *p = row + 4 * C * lane;
        if (C == 4) {
            uint4 t = ldg_nc_v4(p);
            v[0] = t.x; v[1 % C] = t.y; v[2 % C] = t.z; v[3 % C] = t.w;
        } else if (C == 2) {
            uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
            v[0] = t.x; v[1 % C] = t.y;
        } else {
#pragma unroll
            for (int j = 0; j < C; ++j) v[j] = __ldg(reinterpret_cast<const uint32_t *>(p) + j);
        }
    } else {
        const uint8_t *p = row + 4 * C * lane;
#pragma unroll
        for (int j = 0; j < 4 * C; ++j)
            if (j < npx * C) v[j >> 2] |= (uint32_t)__ldg(p + j) << (8 * (j & 3));
    }
}

// Residual bytes of the lane's four pixels.  FLP0 §2: optional subtract-green, then pred = left
// pixel; at x == 0 the pixel above (lane 0 passes it, transformed, in `up`); at (0,0) zero.
// Must be called by all 32 lanes.
template <int C, bool SG>
__device__ __forceinline__ void lane_residuals(uint32_t (&v)[C], uint32_t up, int lane, uint32_t (&res)[C]) {
    if (SG && C == 4) {
#pragma unroll
        for (int j = 0; j < C; ++j) v[j] = subgreen4(v[j]);
    } else if (SG && C == 3) {  // 12 bytes = 4 whole pixels: word j starts at channel phase j, no lane crossing
        const uint32_t a = v[0], b = v[1 % C], c = v[2 % C];
        v[0] = subgreen3(0u, a, b, 0);
        v[1 % C] = subgreen3(a, b, c, 1);
        v[2 % C] = subgreen3(b, c, 0u, 2);
    }
    constexpr int sh = 8 * (4 - C);
    uint32_t pl = __shfl_up_sync(0xFFFFFFFFu, v[C - 1], 1);
    if (lane == 0) pl = up << sh;
    res[0] = __vsub4(v[0], __funnelshift_r(pl, v[0], sh));
#pragma unroll
    for (int j = 1; j < C; ++j) res[j] = __vsub4(v[j], __funnelshift_r(v[j - 1], v[j], sh));
}

// Transformed pixel above the block row's first pixel, packed in the low C bytes (lane 0 only).
// word_ok: the image is 16-byte aligned, so an RGBA / GA pixel can be read as one word / halfword.
template <int C, bool SG>
__device__ __forceinline__ uint32_t up_pixel(const uint8_t *row, uint64_t pitch, bool word_ok) {
    const uint8_t *u = row - pitch;
    uint32_t v;
    if (C == 4 && word_ok) v = __ldg(reinterpret_cast<const uint32_t *>(u));
    else if (C == 2 && word_ok) v = __ldg(reinterpret_cast<const uint16_t *>(u));
    else {
        v = __ldg(u);
        if (C > 1) v |= (uint32_t)__ldg(u + 1) << 8;
        if (C > 2) v |= (uint32_t)__ldg(u + 2) << 16;
        if (C > 3) v |= (uint32_t)__ldg(u + 3) << 24;
    }
    if (SG && C >= 3) {
        const uint32_t g = (v >> 8) & 0xFFu;
        v = __vsub4(v, g | (g << 16));
    }
    return v;
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
// tensor_map: a CUtensorMap for the TMA tile loads (api.cu: make_load_map), or nullptr for direct loads
void launch_histograms(const uint8_t *d_pixels, const Geo &g, uint16_t *d_hist, uint32_t *d_resid, uint2 *d_flat,
                       const void *tensor_map, cudaStream_t s);
void launch_tables(const uint16_t *d_hist, uint64_t nblocks, uint16_t *d_table, uint32_t *d_bits, cudaStream_t s);
void launch_tables_cta(const uint16_t *d_hist, uint64_t nblocks, uint16_t *d_table, uint32_t *d_bits, cudaStream_t s);
int slots_max_resident_ctas();  // how many k_slots CTAs this device keeps resident at once (occupancy API)
bool launch_slots(const Geo &g, const uint32_t *d_bits, unsigned long long *d_dirE, unsigned long long *d_status,
                  uint32_t epoch, uint64_t capacity_words, uint32_t *d_err, uint32_t *d_streams,
                  unsigned long long *d_offsets, int max_grid, cudaStream_t s);
// d_status / d_ticket / ticket_base / epoch: look-back state, used by FLIC_FLAG_EXACT only (the launch draws n * nb tickets)
void launch_pack(const uint32_t *d_resid, const Geo &g, const uint16_t *d_table, const uint2 *d_flat, uint32_t *d_streams,
                 uint64_t capacity_words, unsigned long long *d_dirE, uint32_t *d_err, unsigned long long *d_status,
                 unsigned long long *d_ticket, unsigned long long ticket_base, uint32_t epoch, cudaStream_t s,
                 const unsigned long long *d_part_base = nullptr, uint32_t part_hdr_words = 0);
void launch_finalize(const Geo &g, const unsigned long long *d_dirE, uint32_t *d_streams,
                     uint64_t capacity_words, unsigned long long *d_offsets, uint32_t *d_err,
                     cudaStream_t s);
// fused single-pass encoder; returns the grid size (every CTA draws one ticket beyond the last block)
unsigned launch_encode_fused(const uint8_t *d_pixels, const Geo &g, uint32_t *d_streams, uint64_t capacity_words,
                             unsigned long long *d_dirE, unsigned long long *d_status, unsigned long long *d_ticket,
                             unsigned long long ticket_base, uint32_t epoch, uint32_t *d_err, unsigned long long *d_phase_clk,
                             cudaStream_t s);
// tensor_map: a 128-byte CUtensorMap over the pixel buffer (api.cu: make_pixel_map), or nullptr
void launch_decode(const uint32_t *d_streams, const unsigned long long *d_offsets, const Geo &g,
                   uint8_t *d_pixels, uint32_t *d_err, const void *tensor_map, cudaStream_t s);
void launch_decode_one(const uint32_t *d_streams, const unsigned long long *d_offsets, const Geo &g,
                       uint8_t *d_pixels, uint32_t *d_err, unsigned long long *d_phase_clk, cudaStream_t s);

// block-row splice on the device (splice.cu)
struct SpliceParts {
    uint32_t k;
    uint32_t first_block[FLIC_MAX_PARTS + 1];  // first block of each part in the spliced directory (+ total)
    uint32_t base_words[FLIC_MAX_PARTS + 1];   // payload words before each part (+ total)
};
void launch_splice_finish(uint32_t *d_out, const SpliceParts &sp, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                          cudaStream_t s);
void launch_split_finish(uint32_t *d_part, uint32_t nb, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, cudaStream_t s);
// peer-memory split (flic_encode_emit_device / flic_splice_header_device / flic_pull_part_device)
void launch_part_words(const unsigned long long *d_dirE, uint32_t part_blocks, unsigned long long *d_part_words_out, cudaStream_t s);
void launch_part_directory(const unsigned long long *d_dirE, uint32_t part_blocks, const unsigned long long *d_base_words,
                           uint32_t *d_dir_out, cudaStream_t s);
void launch_splice_header(uint32_t *d_out, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, uint32_t nb,
                          const unsigned long long *d_total_words, uint64_t capacity_words, uint32_t *d_err, cudaStream_t s);
void launch_pull_part(const uint32_t *d_stream, uint64_t stream_words, uint32_t total_blocks, uint32_t first_block, uint32_t part_blocks,
                      uint32_t *d_part, uint64_t capacity_words, unsigned long long *d_part_bytes, uint32_t *d_err, cudaStream_t s);

}  // namespace flic
