// decode.cu — FLP0 decode kernel (sm_100a).
//
// One WARP per block, one LANE per row sub-stream (the format stores 32 word-
// aligned row streams per block precisely so that a warp has 32 independent
// bit-serial decodes in flight).  Per warp: read the 128-byte length table,
// rebuild canonical codes with a packed-counter warp scan, fill a 2^kL-entry
// shared-memory LUT (cooperatively for short codes, per lane for long ones),
// exclusive-scan the row word counts into per-lane stream offsets, then every
// lane runs a branch-light LUT decode with a 64-bit MSB-first bit buffer
// (one refill per 32/kL symbols), undoes the left predictor in registers and
// emits 16-byte vector stores.  Column 0 is a byte-wise prefix sum down the
// rows, done as a warp scan.
//
// Format: DESIGN.md §FLP0 (provisional; not the reference's bitstream).
#include "common.cuh"

namespace flic {

constexpr int kDecWarps = 4;
// a one-word refill guarantees 32 buffered bits: that is floor(32 / kL) whole symbols
constexpr int kSymsPerRefill = 32 / kL;

// predicated 32-bit read-only load: `old` is kept when pred is false (no branch, no access)
__device__ __forceinline__ uint32_t ldg_if(const uint32_t *p, bool pred, uint32_t old) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
        : "+r"(old)
        : "l"(p), "r"((uint32_t)pred));
    return old;
}

// MSB-first reader over one row sub-stream laid out per FLP0 §6: words k < minw sit in the
// block's interleaved region (word k of row r at k*stride + r), the rest in the row's tail.
// Two stream words are always queued in registers (q0 = word k, q1 = word k+1): a word is
// requested four symbols before it is used, which rides out an L1 miss; refill() is
// straight-line code with one predicated load.
struct BitReader {
    const uint32_t *blk;          // block payload (warp-uniform)
    uint32_t ib, stride, tb;      // element offsets: interleaved base, stride, tail base (pre-biased by -minw)
    uint32_t minw, words, k;      // k = index of the word held in q0
    uint32_t q0, q1, n;
    unsigned long long buf;
    __device__ __forceinline__ uint32_t eo(uint32_t i) const { return i < minw ? ib + i * stride : tb + i; }
    __device__ __forceinline__ void init(const uint32_t *b, uint32_t ib_, uint32_t stride_, uint32_t tb_,
                                         uint32_t minw_, uint32_t words_) {
        blk = b; ib = ib_; stride = stride_; tb = tb_; minw = minw_; words = words_;
        k = 0; n = 0; buf = 0;
        q0 = ldg_if(blk + eo(0), 0u < words, 0u);
        q1 = ldg_if(blk + eo(1), 1u < words, 0u);
    }
    // afterwards n >= 32 (exactly 32 when the buffer had run dry), i.e. kSymsPerRefill = 32 / kL
    // symbols are always buffered; the decoder refills that often (every 3rd symbol at kL = 10)
    __device__ __forceinline__ void refill() {
        const bool take = n <= 32u;
        const unsigned long long add = ((unsigned long long)q0 << 32) >> (n & 63u);
        buf |= take ? add : 0ull;
        n += take ? 32u : 0u;
        k += take ? 1u : 0u;
        q0 = take ? q1 : q0;
        q1 = ldg_if(blk + eo(k + 1u), take && k + 1u < words, q1);
    }
    __device__ __forceinline__ uint32_t get(const uint16_t *lut) {
        uint32_t e = lut[(uint32_t)(buf >> (64 - kL))];
        uint32_t l = e >> 8;
        buf <<= l;
        n -= l;
        return e & 0xFFu;
    }
};

__device__ __forceinline__ uint64_t shfl_up64d(uint64_t v, int d) {
    uint32_t lo = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)v, d);
    uint32_t hi = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), d);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t cntd_get(uint64_t a, uint64_t b, uint32_t l) {
    return (uint32_t)((l <= 6 ? a >> (9 * (l - 1)) : b >> (9 * (l - 7))) & 511u);
}
__device__ __forceinline__ void cntd_add(uint64_t &a, uint64_t &b, uint32_t l) {
    if (l >= 1 && l <= 6) a += 1ull << (9 * (l - 1));
    else if (l >= 7 && l <= 12) b += 1ull << (9 * (l - 7));
}

// Builds the warp's LUT from the block's 32 nibble words. Returns false on a malformed table.
__device__ bool build_lut(uint16_t *lut, uint32_t nibw, int lane) {
    {   // entries not covered by any code decode as (symbol 0, 0 bits): loops stay bounded
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *l4 = reinterpret_cast<uint4 *>(lut);
        for (int i = lane; i < kLutSize * 2 / 16; i += 32) l4[i] = z;
    }
    uint32_t l8[8];
    uint64_t ca = 0, cb = 0;
    bool bad = false;
    int sole = -1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        l8[k] = (nibw >> (4 * k)) & 15u;
        if (l8[k] == kLenSole) sole = 8 * lane + k;
        else if (l8[k] > (uint32_t)kL) bad = true;
        cntd_add(ca, cb, l8[k]);
    }
    uint64_t ia = ca, ib = cb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t ta = shfl_up64d(ia, d), tb = shfl_up64d(ib, d);
        if (lane >= d) { ia += ta; ib += tb; }
    }
    uint64_t ea = ia - ca, eb = ib - cb;
    // totals per length live in lane 31's inclusive counters
    uint64_t ta, tb;
    {
        uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)ia, 31), hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(ia >> 32), 31);
        ta = ((uint64_t)hi << 32) | lo;
        lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)ib, 31); hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(ib >> 32), 31);
        tb = ((uint64_t)hi << 32) | lo;
    }
    __syncwarp();

    uint32_t solem = __ballot_sync(0xFFFFFFFFu, sole >= 0);
    if (solem) {
        // one symbol, zero-length code: every LUT entry yields it and consumes nothing
        uint32_t sym = (uint32_t)__shfl_sync(0xFFFFFFFFu, sole, __ffs(solem) - 1);
        uint32_t *l32 = reinterpret_cast<uint32_t *>(lut);
        for (int i = lane; i < kLutSize / 2; i += 32) l32[i] = sym | (sym << 16);
        __syncwarp();
        return !__any_sync(0xFFFFFFFFu, bad);
    }

    // first canonical code of each length, computed redundantly in registers
    uint32_t start[8], span[8], ent[8];
    {
        uint32_t next = 0, prevnum = 0;
        uint32_t nextl[kL + 1];
        nextl[0] = 0;
#pragma unroll
        for (int l = 1; l <= kL; ++l) {
            next = (next + prevnum) << 1;
            nextl[l] = next;
            prevnum = cntd_get(ta, tb, l);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t l = l8[k];
            start[k] = 0; span[k] = 0; ent[k] = 0;
            if (l >= 1 && l <= (uint32_t)kL) {
                uint32_t nl = 0;
#pragma unroll
                for (int t = 1; t <= kL; ++t)
                    if ((uint32_t)t == l) nl = nextl[t];
                uint32_t code = nl + cntd_get(ea, eb, l);
                cntd_add(ea, eb, l);
                start[k] = code << (kL - l);
                span[k] = 1u << (kL - l);
                ent[k] = (uint32_t)(8 * lane + k) | (l << 8);
                if (start[k] + span[k] > (uint32_t)kLutSize) { bad = true; span[k] = 0; }
            }
        }
    }
    // long codes (span <= 16 entries): the owning lane fills them
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (span[k] >= 1 && span[k] <= 16) {
            for (uint32_t i = 0; i < span[k]; ++i) lut[start[k] + i] = (uint16_t)ent[k];
        }
    }
    // short codes (span >= 32): the whole warp fills each span with 32-bit stores
    uint32_t *l32 = reinterpret_cast<uint32_t *>(lut);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t m = __ballot_sync(0xFFFFFFFFu, span[k] >= 32);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            uint32_t st = __shfl_sync(0xFFFFFFFFu, start[k], src);
            uint32_t sp = __shfl_sync(0xFFFFFFFFu, span[k], src);
            uint32_t en = __shfl_sync(0xFFFFFFFFu, ent[k], src);
            for (uint32_t i = lane; i < sp / 2; i += 32) l32[st / 2 + i] = en | (en << 16);
        }
    }
    __syncwarp();
    return !__any_sync(0xFFFFFFFFu, bad);
}

template <int C>
__device__ __forceinline__ uint32_t decode_pixel(BitReader &br, const uint16_t *lut, int &phase) {
    uint32_t r = 0;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        if (phase == 0) br.refill();
        phase = phase == kSymsPerRefill - 1 ? 0 : phase + 1;
        r |= br.get(lut) << (8 * ch);
    }
    return r;
}

template <int C>
__device__ __forceinline__ uint32_t untransform(uint32_t t, bool sg) {
    if (C >= 3 && sg) t = addgreen4(t);
    return t;
}

template <int C>
__device__ __forceinline__ void store_bytes(uint8_t *dst, uint32_t px) {
#pragma unroll
    for (int ch = 0; ch < C; ++ch) dst[ch] = (uint8_t)(px >> (8 * ch));
}

// U pixels fill a whole number of 16-byte chunks: C=1:16, 2:8, 3:16, 4:4
template <int C> struct Unroll { static constexpr int U = (C == 4) ? 4 : (C == 2 ? 8 : 16); };

template <int C>
__device__ void decode_rows(BitReader &br, const uint16_t *lut, uint8_t *dst, int bwa, bool active, bool sg,
                            bool aligned, int lane) {
    constexpr int U = Unroll<C>::U;
    constexpr int W = U * C / 4;  // words per chunk
    constexpr uint32_t cmask = C == 4 ? 0xFFFFFFFFu : ((1u << (8 * (C & 3))) - 1u);
    // column 0: residual against the pixel above == byte-wise prefix sum down the rows
    int phase = 0;
    uint32_t cur = active ? decode_pixel<C>(br, lut, phase) : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, cur, d);
        if (lane >= d) cur = __vadd4(cur, t);
    }
    if (!active) return;

    int x = 0;
    for (; x + U <= bwa; x += U) {
        uint32_t o[W];
#pragma unroll
        for (int i = 0; i < W; ++i) o[i] = 0;
        phase = 0;  // chunk boundary: always refill first
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u > 0 || x > 0) cur = __vadd4(cur, decode_pixel<C>(br, lut, phase));
            uint32_t px = untransform<C>(cur, sg) & cmask;
            const int bp = u * C;
            o[bp >> 2] |= px << (8 * (bp & 3));
            if ((bp & 3) + C > 4) o[(bp >> 2) + 1] |= px >> (8 * (4 - (bp & 3)));
        }
        uint8_t *d = dst + (size_t)x * C;
        if (aligned) {
#pragma unroll
            for (int i = 0; i < W; i += 4)
                *reinterpret_cast<uint4 *>(d + 4 * i) = make_uint4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < W * 4; ++i) d[i] = (uint8_t)(o[i >> 2] >> (8 * (i & 3)));
        }
    }
    // ragged right edge of the image
    for (; x < bwa; ++x) {
        phase = 0;
        if (x > 0) cur = __vadd4(cur, decode_pixel<C>(br, lut, phase));
        store_bytes<C>(dst + (size_t)x * C, untransform<C>(cur, sg));
    }
}

__global__ void __launch_bounds__(kDecWarps * 32, 10) k_decode(const uint32_t *__restrict__ streams,
                                                          const unsigned long long *__restrict__ offsets, Geo g,
                                                          uint8_t *__restrict__ pixels, uint32_t *err) {
    __shared__ __align__(16) uint16_t luts[kDecWarps][kLutSize];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t gb = (uint64_t)blockIdx.x * kDecWarps + warp;
    if (gb >= (uint64_t)g.n * g.nb) return;
    const BlockPos p = block_pos(g, gb);
    uint16_t *lut = luts[warp];

    const unsigned long long sbeg = offsets[p.img], send = offsets[p.img + 1];
    const uint32_t *sw = streams + (sbeg >> 2);
    const uint64_t swords = (send - sbeg) >> 2;
    const uint64_t fixed = kHdrWords + (uint64_t)g.nb + 1;
    bool ok = (sbeg & 3) == 0 && send >= sbeg && swords >= fixed;
    uint32_t off = 0, end = 0, pw = 0;
    if (ok) {
        pw = sw[6];
        off = sw[kHdrWords + p.b];
        end = sw[kHdrWords + p.b + 1];
        ok = sw[0] == kMagic && (sw[1] & 0xFFFFu) == 2u && sw[2] == g.w && sw[3] == g.h && sw[5] == g.nb && fixed + pw <= swords &&
             off <= end && end <= pw && end - off >= (uint32_t)kBlkHdrWords;
    }
    if (!ok) {
        if (lane == 0) atomicOr(err, kErrFormat);
        return;
    }
    const uint32_t *blk = sw + fixed + off;

    ok = build_lut(lut, __ldg(blk + lane), lane);

    const bool active = lane < (int)p.bha;
    uint32_t rc = (__ldg(blk + 32 + (lane >> 1)) >> (16 * (lane & 1))) & 0xFFFFu;
    uint32_t incl = warp_incl_scan(rc, lane);
    uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t minw = active ? rc : 0xFFFFFFFFu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) minw = min(minw, __shfl_xor_sync(0xFFFFFFFFu, minw, d));
    ok = ok && (uint32_t)kBlkHdrWords + total <= end - off && !__any_sync(0xFFFFFFFFu, !active && rc != 0);
    if (!ok) {
        if (lane == 0) atomicOr(err, kErrFormat);
        return;
    }
    BitReader br;
    br.init(blk, kBlkHdrWords + lane, p.bha, kBlkHdrWords + minw * p.bha + (incl - rc) - lane * minw - minw, minw,
            active ? rc : 0u);
    uint8_t *dst = pixels + (uint64_t)p.img * g.img_stride + (uint64_t)(p.y0 + lane) * g.pitch +
                   (uint64_t)p.x0 * g.c;
    const bool sg = (g.flags & FLIC_FLAG_SUBGREEN) != 0;
    const bool aligned = g.aligned16 != 0;
    switch (g.c) {
        case 1: decode_rows<1>(br, lut, dst, (int)p.bwa, active, sg, aligned, lane); break;
        case 2: decode_rows<2>(br, lut, dst, (int)p.bwa, active, sg, aligned, lane); break;
        case 3: decode_rows<3>(br, lut, dst, (int)p.bwa, active, sg, aligned, lane); break;
        default: decode_rows<4>(br, lut, dst, (int)p.bwa, active, sg, aligned, lane); break;
    }
}

void launch_decode(const uint32_t *d_streams, const unsigned long long *d_offsets, const Geo &g,
                   uint8_t *d_pixels, uint32_t *d_err, cudaStream_t s) {
    uint64_t total = (uint64_t)g.n * g.nb;
    unsigned grid = (unsigned)((total + kDecWarps - 1) / kDecWarps);
    k_decode<<<grid, kDecWarps * 32, 0, s>>>(d_streams, d_offsets, g, d_pixels, d_err);
}

}  // namespace flic
