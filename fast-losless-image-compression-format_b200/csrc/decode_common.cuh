// decode_common.cuh — pieces shared by the two decode kernels (decode.cu: one lane per row sub-stream;
// decode_one.cu: self-synchronising decode of one bit stream per block): the shared-memory Huffman LUT
// builder and the pixel pack helpers.  Format: DESIGN.md §FLP0 (provisional; not the reference's bitstream).
#pragma once

#include "common.cuh"

namespace flic {

// Per-warp scratch of the LUT build.
struct LutScratch {
    uint16_t ent[256 + 2];  // LUT entries in canonical (length, symbol) order, then an all-zero terminator
    uint32_t starts[32];    // bit v set: a code's span begins at LUT index v
    uint32_t cnt[16];       // symbols per code length (running during the rank rounds)
    uint32_t base[16];      // first LUT index of each length class
    uint32_t below[16];     // symbols with a shorter code
};

// Builds the warp's LUT (2^kL entries of  len | symbol << 8) from the block's 32 nibble words.
// Returns false on a malformed table.
//
// Canonical codes ordered by (length, symbol) tile the LUT with one contiguous span per symbol.  Eight
// rounds give every symbol its rank within its length class (round r: lane L holds symbol 32r + L, the
// lanes of equal length find each other with MATCH.ANY, a per-length running count lives in shared
// memory); a scan over the ten lengths gives each class's first index; every symbol then drops its
// entry at its canonical rank and sets the bit of its span's first index.  Finally each lane fills its
// own 32 consecutive LUT entries — "how many spans have started up to here" is a popcount of the start
// bits — with four 16-byte stores: no loops whose trip counts depend on the code, no conflicts.
static __device__ bool build_lut(uint16_t *lut, LutScratch &sc, uint32_t nibw, int lane) {
    if (lane < 16) { sc.cnt[lane] = 0; }
    sc.starts[lane] = 0;
    __syncwarp();
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t len[8], rank[8];
    bool bad = false;
    int sole = -1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {  // symbol 32r + lane: nibble (lane & 7) of word 4r + (lane >> 3)
        const uint32_t w = __shfl_sync(0xFFFFFFFFu, nibw, 4 * r + (lane >> 3));
        const uint32_t l = (w >> (4 * (lane & 7))) & 15u;
        len[r] = l;
        if (l == kLenSole) sole = 32 * r + lane;
        else if (l > (uint32_t)kL) bad = true;
        const uint32_t m = __match_any_sync(0xFFFFFFFFu, l);
        const uint32_t before = sc.cnt[l];
        rank[r] = before + __popc(m & lt);
        __syncwarp();
        if ((m & lt) == 0) sc.cnt[l] = before + __popc(m);  // the group's first lane
        __syncwarp();
    }
    const uint32_t solem = __ballot_sync(0xFFFFFFFFu, sole >= 0);
    if (solem) {
        // one symbol, zero-length code: every LUT entry yields it and consumes nothing
        const uint32_t sym = (uint32_t)__shfl_sync(0xFFFFFFFFu, sole, __ffs(solem) - 1);
        uint4 *l4 = reinterpret_cast<uint4 *>(lut);
        const uint32_t e2 = (sym << 8) | (sym << 24);
#pragma unroll
        for (int i = 0; i < 4; ++i) l4[4 * lane + i] = make_uint4(e2, e2, e2, e2);
        __syncwarp();
        return !__any_sync(0xFFFFFFFFu, bad);
    }
    // length classes: lane l (1..kL) owns class l
    {
        const uint32_t n = (lane >= 1 && lane <= kL) ? sc.cnt[lane] : 0u;
        const uint32_t span = n << ((kL - lane) & 31);
        uint32_t is = span, in = n;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const uint32_t ts = __shfl_up_sync(0xFFFFFFFFu, is, d), tn = __shfl_up_sync(0xFFFFFFFFu, in, d);
            if (lane >= d) { is += ts; in += tn; }
        }
        if (lane < 16) { sc.base[lane] = is - span; sc.below[lane] = in - n; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, is, kL), nsym = __shfl_sync(0xFFFFFFFFu, in, kL);
        if (total > (uint32_t)kLutSize) bad = true;  // over-subscribed lengths
        if (lane == 0) {
            sc.ent[nsym] = 0;  // terminator: indices past the last span decode as (symbol 0, 0 bits)
            if (total < (uint32_t)kLutSize) sc.starts[total >> 5] = 1u << (total & 31);
        }
    }
    __syncwarp();
    if (__any_sync(0xFFFFFFFFu, bad)) return false;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t l = len[r];
        if (l) {
            const uint32_t start = sc.base[l] + (rank[r] << (kL - l));
            sc.ent[sc.below[l] + rank[r]] = (uint16_t)(l | ((uint32_t)(32 * r + lane) << 8));
            atomicOr(&sc.starts[start >> 5], 1u << (start & 31));
        }
    }
    __syncwarp();
    // lane L fills LUT entries 32L .. 32L+31
    const uint32_t flags = sc.starts[lane];
    uint32_t upto = __popc(flags);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, upto, d);
        if (lane >= d) upto += t;
    }
    const uint32_t first = upto - __popc(flags) - 1u;  // index of the span covering entry 32L - 1 (or -1)
    uint4 *l4 = reinterpret_cast<uint4 *>(lut) + 4 * lane;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = 8 * q + 2 * j;
            const uint32_t e0 = sc.ent[first + __popc(flags & ((2u << i) - 1u))];
            const uint32_t e1 = sc.ent[first + __popc(flags & (i + 1 == 31 ? 0xFFFFFFFFu : ((2u << (i + 1)) - 1u)))];
            w[j] = e0 | (e1 << 16);
        }
        l4[q] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncwarp();
    return true;
}

// four pixels (C valid low bytes each) -> C output words
template <int C>
__device__ __forceinline__ void pack4(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t *o) {
    if (C == 4) { o[0] = p0; o[1] = p1; o[2] = p2; o[3] = p3; }
    else if (C == 3) {
        o[0] = __byte_perm(p0, p1, 0x4210);
        o[1] = __byte_perm(p1, p2, 0x5421);
        o[2] = __byte_perm(p2, p3, 0x6542);
    } else if (C == 2) {
        o[0] = __byte_perm(p0, p1, 0x5410);
        o[1] = __byte_perm(p2, p3, 0x5410);
    } else {
        o[0] = __byte_perm(__byte_perm(p0, p1, 0x0040), __byte_perm(p2, p3, 0x0040), 0x5410);
    }
}

template <int C>
__device__ __forceinline__ void store_bytes(uint8_t *dst, uint32_t px) {
#pragma unroll
    for (int ch = 0; ch < C; ++ch) dst[ch] = (uint8_t)(px >> (8 * ch));
}

}  // namespace flic
