// decode.cu — FLP0 decode kernel (sm_100a).
//
// One WARP per block, one LANE per row sub-stream (the format stores 32 word-
// aligned row streams per block precisely so that a warp has 32 independent
// bit-serial decodes in flight).  Per warp: read the 128-byte length table,
// rank the symbols within their length classes (MATCH.ANY rounds), fill the
// 2^kL-entry shared-memory LUT (every lane its own 32 entries), exclusive-scan
// the row word counts into per-lane stream offsets, then every lane runs a LUT
// decode whose per-symbol work is shaped around the ALU pipe (see BitReader),
// undoes the left predictor in 16-bit-lane running sums and hands RGBA / RGB pixels
// to the TMA unit in 32 x 64 B / 32 x 96 B tiles (other layouts: 256-/128-bit
// stores).  Column 0 is a byte-wise prefix sum down the rows, done as a warp scan.
// The ONE_STREAM layout has its own decoder (decode_one.cu).
//
// Format: DESIGN.md §FLP0 (provisional; not the reference's bitstream).
#include <cstring>

#include <cuda.h>  // CUtensorMap (type only; the driver entry point is resolved in api.cu)

#include "decode_common.cuh"

namespace flic {

constexpr int kDecWarps = 4;
constexpr int kLutPad = 5 * 1024;  // offset of the LUTs in k_decode's shared struct: >= scratch + records, and 1 KB (mod 2 KB)
static_assert(sizeof(LutScratch) * kDecWarps + 32 * kDecWarps <= kLutPad && kLutPad % 2048 == 1024, "k_decode shared layout");
// stream lines are pulled into L1 this many chunks of average consumption ahead (measured: 0 -> 2 is -7 %)
constexpr uint32_t kPrefetchChunks = 2;
// ... and the head of a block is pulled towards L2 while its LUT is built.  2 KB: more (the whole block)
// gets evicted by the output stream before it is used and is read from DRAM twice (measured: DRAM reads
// 0.96 GB with 0-2 KB, 1.06 GB with 4 KB, 1.46 GB with the whole block; time best at 2 KB)
constexpr uint32_t kPrefetchL2Words = 512;
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
//
// The loop is built around what the ALU pipe (LOP3/SHF/IADD, half rate on sm_100) has to do per
// symbol, which is what bounds this kernel:
//   * the 64-bit bit buffer is two registers, valid bits at the top of hi, zeros below them;
//   * a LUT entry is  len | symbol << 8 : the funnel shifts that consume a symbol read their
//     count from the entry's low 5 bits (wrap mode), so the length is never extracted;
//   * `cn` counts buffered bits in its low byte and is updated as cn -= entry (the symbol byte
//     only disturbs bits 8 and up); "needs a refill" is a single bit test, cn & 32 == 0;
//   * the LUT address is (hi >> 21) & 0x7FE | lut: the shift is issued as IMAD.HI on the FMA
//     pipe, the mask-and-base is one LOP3 (the warp's LUT is 2 KB-aligned in shared memory);
//   * a refill is straight-line predicated code; one stream word is always queued in `wq`.
struct BitReader {
    uint32_t hi, lo;  // bit buffer
    uint32_t cn;      // low byte: number of buffered bits (0..63)
    uint32_t wq;      // next stream word (index k - 1), already loaded
    uint32_t k;       // index of the word after wq
};

struct RowStream {
    const uint32_t *blk;     // block payload (warp-uniform)
    const char *lane0;       // the lane's word 0 in the interleaved region: word k is at lane0 + k * stride4
    uint32_t stride4;        // stride in bytes
    uint32_t ib, stride, tb; // element offsets: interleaved base, stride, tail base (pre-biased by -minw)
    uint32_t minw, words;
    __device__ __forceinline__ uint32_t eo(uint32_t i) const { return i < minw ? ib + i * stride : tb + i; }
};

__device__ __forceinline__ void reader_init(BitReader &r, const RowStream &rs) {
    r.hi = r.lo = 0; r.cn = 0; r.k = 1;
    r.wq = ldg_if(rs.blk + rs.eo(0), 0u < rs.words, 0u);
}

// After this at least 32 bits are buffered, i.e. kSymsPerRefill = 32 / kL whole symbols.
// kFast: every word this chunk can touch lies in the interleaved region (checked by the caller
// with one vote per chunk), so the address is one IMAD and there is no bounds test.
template <bool kFast>
__device__ __forceinline__ void refill(BitReader &r, const RowStream &rs) {
    // Straight-line, predicated (hand-scheduled in PTX so that it stays six ALU operations):
    //   take = (cn & 32) == 0          fewer than 32 bits buffered, so lo is empty
    //   hi  |= wq >> n;  lo = wq << (32 - n)   (funnel shifts in wrap mode read n from cn's low 5 bits;
    //                                           n == 0 gives lo = 0)
    //   cn  += 32;  wq = next word;  k += 1
    if (kFast) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(rs.lane0 + (uint64_t)r.k * rs.stride4);  // one IMAD.WIDE
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
            "and.b32 t, %2, 32;\n\t"
            "setp.eq.u32 p, t, 0;\n\t"
            "@p shf.r.wrap.b32 t, %3, 0, %2;\n\t"
            "@p or.b32 %0, %0, t;\n\t"
            "@p shf.r.wrap.b32 %1, 0, %3, %2;\n\t"
            "@p add.u32 %2, %2, 32;\n\t"
            "@p ld.global.nc.u32 %3, [%5];\n\t"
            "@p add.u32 %4, %4, 1;\n\t}"
            : "+r"(r.hi), "+r"(r.lo), "+r"(r.cn), "+r"(r.wq), "+r"(r.k)
            : "l"(p));
    } else {
        const uint32_t *p = rs.blk + rs.eo(r.k);
        const uint32_t in = r.k < rs.words;  // past the row's last word the stream reads as zeros
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .b32 t;\n\t"
            "and.b32 t, %2, 32;\n\t"
            "setp.eq.u32 p, t, 0;\n\t"
            "setp.ne.and.u32 q, %6, 0, p;\n\t"
            "@p shf.r.wrap.b32 t, %3, 0, %2;\n\t"
            "@p or.b32 %0, %0, t;\n\t"
            "@p shf.r.wrap.b32 %1, 0, %3, %2;\n\t"
            "@p add.u32 %2, %2, 32;\n\t"
            "@p mov.u32 %3, 0;\n\t"
            "@q ld.global.nc.u32 %3, [%5];\n\t"
            "@p add.u32 %4, %4, 1;\n\t}"
            : "+r"(r.hi), "+r"(r.lo), "+r"(r.cn), "+r"(r.wq), "+r"(r.k)
            : "l"(p), "r"(in));
    }
}

// Decodes one symbol; returns the LUT entry (len | sym << 8).  `lut_s` is the shared-window address of
// the warp's LUT, 2 KB-aligned, so (index & 0x7FE) | lut_s is the entry's complete address in ONE LOP3 and the
// load needs no base add (a generic pointer + offset cost an extra IMAD.IADD per symbol); `m2048` is the
// constant 2048 kept opaque (a kernel argument) so that the shift is issued as IMAD.HI.
__device__ __forceinline__ uint32_t get(BitReader &r, uint32_t lut_s, uint32_t m2048) {
    const uint32_t a = (__umulhi(r.hi, m2048) & 0x7FEu) | lut_s;
    uint32_t e;  // (the LUT is complete, and fenced by __syncwarp, before any reader exists: plain asm, free to be scheduled)
    asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(a));
    r.hi = __funnelshift_l(r.lo, r.hi, e);
    r.lo = __funnelshift_l(0u, r.lo, e);
    r.cn -= e;
    return e;
}

// ---- per-lane pixel reconstruction ------------------------------------------------------------
// Left prediction is a running byte-wise sum along the row.  The running values live in two
// registers with 16-bit lanes, A = (ch0, ch2) and B = (ch1, ch3), each value in the HIGH byte of
// its half: adding a raw LUT entry (len | sym << 8) advances the value, and the length lands in
// the low byte, where it is junk that is masked off once per chunk (before it can carry).
struct Acc { uint32_t a, b; };

template <int C>
__device__ __forceinline__ void acc_add(Acc &v, int ch, uint32_t e) {
    if (ch == 0) v.a += e;
    else if (ch == 1) v.b += e;
    else if (ch == 2) v.a = e * 65536u + v.a;
    else v.b = e * 65536u + v.b;
}
__device__ __forceinline__ void acc_clean(Acc &v) { v.a &= 0xFF00FF00u; v.b &= 0xFF00FF00u; }

// accumulators holding packed (transformed) pixel t
template <int C>
__device__ __forceinline__ Acc acc_from(uint32_t t) {
    Acc v;
    v.a = __byte_perm(t, 0u, 0x2404);  // [0, t.b0, 0, t.b2]
    v.b = __byte_perm(t, 0u, 0x3414);  // [0, t.b1, 0, t.b3]
    return v;
}
// the C bytes of the current pixel in the low bytes (colour transform undone); bytes above C are junk
template <int C, bool SG>
__device__ __forceinline__ uint32_t acc_pixel(const Acc &v) {
    uint32_t a = v.a;
    if (SG && C >= 3) a += __byte_perm(v.b, 0u, 0x1414);  // R += G, B += G (carries only reach junk bytes)
    return __byte_perm(a, v.b, 0x7351);                   // [a.b1, b.b1, a.b3, b.b3]
}

// U pixels per chunk fill whole 16-byte stores (two for RGBA: a full 32-byte sector).
// FM: compile-time mask of flat channels (FLP0 §2b) — they carry no symbols; CS = symbols per pixel.
template <int C, int FM> struct Chunk {
    static constexpr int U = (C == 4) ? 8 : (C == 2 ? 8 : 16);
    static constexpr int W = U * C / 4;  // words per chunk
    static constexpr int CS = C - ((FM & 1) + ((FM >> 1) & 1) + ((FM >> 2) & 1) + ((FM >> 3) & 1));
    static constexpr int kMaxRefills = (U * CS * kL + 31) / 32 + 1;  // words a chunk can request
};

// One chunk of U pixels: U*CS symbols with a refill check before every kSymsPerRefill-th.
template <int C, bool SG, int FM, bool kFast>
__device__ __forceinline__ void decode_chunk(BitReader &br, const RowStream &rs, Acc &acc, uint32_t lut_s,
                                             uint32_t m2048, uint32_t *o) {
    constexpr int U = Chunk<C, FM>::U;
    int sidx = 0;  // compile-time after unrolling
#pragma unroll
    for (int g = 0; g < U / 4; ++g) {
        uint32_t px[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                if ((FM >> ch) & 1) continue;
                if (sidx % kSymsPerRefill == 0) refill<kFast>(br, rs);
                ++sidx;
                acc_add<C>(acc, ch, get(br, lut_s, m2048));
            }
            px[u] = acc_pixel<C, SG>(acc);
        }
        pack4<C>(px[0], px[1], px[2], px[3], o + g * C);
    }
    acc_clean(acc);
}

// channel `ch` of the running sums := v (flat channels hold their constant for the whole block)
__device__ __forceinline__ void acc_set(Acc &v, int ch, uint32_t val) {
    const uint32_t keep = (ch & 2) ? 0x0000FFFFu : 0xFFFF0000u, put = val << ((ch & 2) ? 24 : 8);
    if (ch & 1) v.b = (v.b & keep) | put;
    else v.a = (v.a & keep) | put;
}

// FM >= 0: the block's flat mask is the compile-time FM (0, or 8 = opaque-alpha RGBA), rows run in unrolled
// chunks.  FM < 0: any other mask, taken from `fmask` at run time — a plain per-pixel loop (rare blocks).
// ---- TMA store path (RGBA, 16-byte-aligned images) ----------------------------------------------
// A lane's 32 B stores each touch their own 128 B line, so a warp-wide store costs the LSU data pipe 32
// wavefronts for 1 KB — and that pipe is what bounds this kernel.  Instead two chunks (16 pixels = 64 B
// per row) are staged in shared memory with conflict-free 16 B stores (4 wavefronts each) and one thread
// hands the 32 x 64 B tile to the TMA unit (cp.async.bulk.tensor, 3-D map {row bytes, rows, images}, so
// tiles clip at the image's right and bottom edges).  The tile uses the 64 B swizzle: 16 B chunk c of row
// r sits at chunk c ^ ((r >> 1) & 3).  Register budget: the loop only carries the tile's shared address
// (0 = path off); the tile's origin and the descriptor address sit in a record right behind the tile and
// are read by lane 0 per flush.
// RGB rows go the same way with 96-byte tile rows (two chunks of 16 pixels) and no swizzle: a lane's 16 B stores at a
// 96-byte pitch are 2-way bank-conflicted, still four times fewer wavefronts than stores straight to global memory.
// kTma selects the variant at compile time: 0 none, 4 RGBA, 3 RGB (each kernel carries only its own staging).
template <int kTma> struct TileGeo {
    static constexpr uint32_t kBytes = kTma == 3 ? 32u * 96u : 32u * 64u;  // a multiple of 1 KB: every tile keeps the swizzle phase
};
struct TileRec { unsigned long long map; uint32_t x0b, y0, img, pad[3]; };  // 32 bytes
static_assert(sizeof(TileRec) == 32, "tile_flush addresses the warp's record as recs + 32 * warp");
__device__ __forceinline__ void tile_put(uint32_t tile, int lane, int half, const uint32_t *o) {
    const uint32_t row = tile + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const uint32_t a = row + ((((uint32_t)(2 * half + c)) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(o[4 * c]), "r"(o[4 * c + 1]), "r"(o[4 * c + 2]),
                     "r"(o[4 * c + 3]) : "memory");
    }
}
// amask: the warp's lanes that own a real row (lanes of a ragged bottom block's missing rows have
// left decode_rows, so a full-mask barrier here would name exited lanes).  Lane 0 always owns a row.
__device__ __forceinline__ void tile_put3(uint32_t tile, int lane, int half, const uint32_t *o) {  // RGB: 48 B per chunk
    const uint32_t a = tile + (uint32_t)lane * 96u + (uint32_t)half * 48u;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a + 16u * c), "r"(o[4 * c]), "r"(o[4 * c + 1]), "r"(o[4 * c + 2]),
                     "r"(o[4 * c + 3]) : "memory");
}
// rec_s: shared address of the warp's TileRec (recomputed per flush from the warp index: no register lives across the loop)
__device__ __forceinline__ void tile_flush(uint32_t tile, uint32_t rec_s, int lane, uint32_t xbyte, uint32_t amask) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
    __syncwarp(amask);
    if (lane == 0) {
        unsigned long long map;
        uint32_t x0b, y0, img;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(map) : "r"(rec_s));
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x0b), "=r"(y0) : "r"(rec_s + 8));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(img) : "r"(rec_s + 16));
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
                     "r"(x0b + xbyte), "r"(y0), "r"(img), "r"(tile) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
__device__ __forceinline__ void tile_wait(int lane, uint32_t amask) {  // the tile may be overwritten once the TMA unit has read it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp(amask);
}

template <int C, bool SG, int FM, int kTma>
__device__ void decode_rows(const RowStream &rs, uint32_t lut_s, uint32_t m2048, uint8_t *dst, int bwa,
                            bool active, int aligned, int lane, uint32_t fmask, uint32_t fvals, uint32_t pfx,
                            uint32_t tma /* the warp's tile, or 0 */, uint32_t recs_s /* the CTA's TileRec array */) {
    constexpr int FMC = FM < 0 ? 0 : FM;
    constexpr int U = Chunk<C, FMC>::U, W = Chunk<C, FMC>::W;
    const uint32_t amask = __ballot_sync(0xFFFFFFFFu, active);
    BitReader br;
    reader_init(br, rs);
    // Column 0 is predicted from the pixel above: its value is a byte-wise prefix sum of the rows'
    // first residuals.  Decode that one pixel on a COPY of the reader, scan, and start the row's
    // running sum from the value above; the main loop then decodes the row from x = 0 uniformly.
    uint32_t first = 0;
    if (active) {
        BitReader t = br;
        Acc z = {0u, 0u};
        int sidx = 0;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            if ((fmask >> ch) & 1u) continue;
            if (sidx % kSymsPerRefill == 0) refill<false>(t, rs);
            ++sidx;
            acc_add<C>(z, ch, get(t, lut_s, m2048));
        }
        first = __byte_perm(z.a, z.b, 0x7351);
        if (C < 4) first &= (1u << (8 * (C & 3))) - 1u;
    }
    uint32_t incl = first;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl = __vadd4(incl, t);
    }
    uint32_t above = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) above = 0;
    if (!active) return;
    Acc acc = acc_from<C>(above);
#pragma unroll
    for (int ch = 0; ch < C; ++ch)
        if ((fmask >> ch) & 1u) acc_set(acc, ch, (fvals >> (8 * ch)) & 0xFFu);

    int x = 0;
    if (FM >= 0) {
        // L1 prefetch distance in stream words: `pfx` chunks' worth of this row's average consumption
        const uint32_t pf = pfx ? (pfx * rs.words * (uint32_t)U) / (uint32_t)bwa + 2u : 0u;
        for (; x + U <= bwa; x += U) {
            uint32_t o[W];
            // fast path while no lane of the warp can leave the interleaved region inside this chunk
            if (__all_sync(amask, br.k + (uint32_t)Chunk<C, FMC>::kMaxRefills <= rs.minw)) {
                if (pf) {  // pull the interleaved line `pf` words ahead into L1: refills then hit it
                    const uint32_t kp = min(br.k + pf, rs.minw - 1u);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(rs.blk + (rs.ib + kp * rs.stride)));
                }
                decode_chunk<C, SG, FMC, true>(br, rs, acc, lut_s, m2048, o);
            }
            else
                decode_chunk<C, SG, FMC, false>(br, rs, acc, lut_s, m2048, o);
            uint8_t *d = dst + (size_t)x * C;
            const int half = (x / U) & 1;
            if (kTma == 4 && C == 4 && tma && (half == 1 || x + 2 * U <= bwa)) {  // chunk pairs go out as one TMA tile
                if (half == 0) { if (x > 0) tile_wait(lane, amask); tile_put(tma, lane, 0, o); }
                else { tile_put(tma, lane, 1, o); tile_flush(tma, recs_s + 32u * (threadIdx.x >> 5), lane, (uint32_t)(x - U) * 4u, amask); }
            } else if (kTma == 3 && C == 3 && tma && (half == 1 || x + 2 * U <= bwa)) {
                if (half == 0) { if (x > 0) tile_wait(lane, amask); tile_put3(tma, lane, 0, o); }
                else { tile_put3(tma, lane, 1, o); tile_flush(tma, recs_s + 32u * (threadIdx.x >> 5), lane, (uint32_t)(x - U) * 3u, amask); }
            } else if (W == 8 && aligned == 2) {  // one 256-bit store: a whole 32-byte sector per lane and half the LSU wavefronts
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                             "r"(o[3]), "r"(o[4 % W]), "r"(o[5 % W]), "r"(o[6 % W]), "r"(o[7 % W])
                             : "memory");
            } else if (aligned) {
#pragma unroll
                for (int i = 0; i < W; i += 4)
                    *reinterpret_cast<uint4 *>(d + 4 * i) = make_uint4(o[i], o[i + 1], o[i + 2], o[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < W * 4; ++i) d[i] = (uint8_t)(o[i >> 2] >> (8 * (i & 3)));
            }
        }
    }
    if (kTma == C && tma && FM >= 0) tile_wait(lane, amask);  // the tile must outlive the TMA unit's read of it
    // ragged right edge of the image, and whole rows of blocks with an uncommon flat mask
    for (; x < bwa; ++x) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            if ((fmask >> ch) & 1u) continue;  // warp-uniform
            refill<false>(br, rs);
            acc_add<C>(acc, ch, get(br, lut_s, m2048));
        }
        store_bytes<C>(dst + (size_t)x * C, acc_pixel<C, SG>(acc));
        acc_clean(acc);
    }
}

// kTma: the RGBA TMA-store variant; the plain variant carries none of that code (its other paths were
// measurably slower with it compiled in).
template <int kTma>
__global__ void __launch_bounds__(kDecWarps * 32, kTma == 3 ? 8 : 10) k_decode(const uint32_t *__restrict__ streams,
                                                          const unsigned long long *__restrict__ offsets, Geo g,
                                                          uint8_t *__restrict__ pixels, uint32_t *err, uint32_t m2048, uint32_t pf, uint32_t pf2,
                                                          const __grid_constant__ CUtensorMap tmap) {
    // ONE struct, so the order is ours.  get() ORs a LUT index into the warp's LUT address, which therefore has to be
    // 2 KB-aligned in the shared WINDOW; static shared memory starts 1 KB into it (the driver's reserved KB), so the LUTs
    // sit at an offset of 1 KB (mod 2 KB) — checked below, a violation is reported, never decoded through.
    struct Smem {
        LutScratch scratch[kDecWarps];
        TileRec recs[kDecWarps];
        uint8_t pad[kLutPad - sizeof(LutScratch) * kDecWarps - sizeof(TileRec) * kDecWarps];
        uint16_t luts[kDecWarps][kLutSize];
        uint8_t tiles[kTma ? kDecWarps : 1][kTma ? TileGeo<kTma>::kBytes : 16];  // TMA store staging, 32 rows x 64 / 96 B per warp
    };
    static_assert(offsetof(Smem, luts) % 2048 == 1024 && offsetof(Smem, tiles) % 1024 == 0, "see above");
    __shared__ __align__(1024) Smem sm;
    auto &luts = sm.luts; auto &tiles = sm.tiles; auto &scratch = sm.scratch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t gb = (uint64_t)blockIdx.x * kDecWarps + warp;
    if (gb >= (uint64_t)g.n * g.nb) return;
    const BlockPos p = block_pos(g, gb);
    uint16_t *lut = luts[warp];
    const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(lut);
    if (lut_s & 2047u) {  // the layout assumption above does not hold on this driver: refuse
        if (lane == 0) atomicOr(err, kErrLayout);
        return;
    }

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
        // the header must describe exactly the geometry this launch decodes with: a stream of another channel
        // count, colour transform or block shape would otherwise decode into plausible garbage
        ok = sw[0] == kMagic && ((sw[1] ^ (kVersion | (g.c << 16) | ((g.flags & 0xFFu) << 24))) & ~((uint32_t)FLIC_FLAG_EXACT << 24)) == 0 &&
             sw[2] == g.w && sw[3] == g.h &&  // (EXACT only changes where the encoder put the blocks: the directory says)
             sw[4] == ((uint32_t)kBW | ((uint32_t)kBH << 16)) && sw[5] == g.nb && sw[7] == (uint32_t)kL && fixed + pw <= swords &&
             off <= end && end <= pw && end - off >= (uint32_t)kBlkHdrWords;
    }
    if (!ok) {
        if (lane == 0) atomicOr(err, kErrFormat);
        return;
    }
    const uint32_t *blk = sw + fixed + off;
    if (pf2) {  // pull the block's head towards L2 now: the LUT build below covers the DRAM latency
        const uint32_t lim = min(end - off, pf2);
        for (uint32_t i = 32u * lane; i < lim; i += 32u * 32u) asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + i));
    }

    ok = build_lut(lut, scratch[warp], __ldg(blk + lane), lane);

    const bool active = lane < (int)p.bha;
    uint32_t rc = (__ldg(blk + 32 + (lane >> 1)) >> (16 * (lane & 1))) & 0xFFFFu;
    uint32_t incl = warp_incl_scan(rc, lane);
    uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t minw = active ? rc : 0xFFFFFFFFu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) minw = min(minw, __shfl_xor_sync(0xFFFFFFFFu, minw, d));
    const uint32_t fmask = __ldg(blk + kFlatWord), fvals = __ldg(blk + kFlatWord + 1);  // FLP0 §2b
    ok = ok && (uint32_t)kBlkHdrWords + total <= end - off && !__any_sync(0xFFFFFFFFu, !active && rc != 0) && (fmask >> g.c) == 0;
    if (!ok) {
        if (lane == 0) atomicOr(err, kErrFormat);
        return;
    }
    RowStream rs;
    rs.blk = blk; rs.ib = kBlkHdrWords + lane; rs.stride = p.bha;
    rs.lane0 = reinterpret_cast<const char *>(blk + rs.ib); rs.stride4 = 4u * p.bha;
    rs.tb = kBlkHdrWords + minw * p.bha + (incl - rc) - lane * minw - minw;
    rs.minw = minw; rs.words = active ? rc : 0u;
    uint8_t *dst = pixels + (uint64_t)p.img * g.img_stride + (uint64_t)(p.y0 + lane) * g.pitch +
                   (uint64_t)p.x0 * g.c;
    const bool sg = (g.flags & FLIC_FLAG_SUBGREEN) != 0 && g.c >= 3;
    const int aligned = g.aligned16 ? (g.aligned32 ? 2 : 1) : 0;  // 0: byte stores, 1: 16-byte, 2: 32-byte
    uint32_t tp = 0;
    if (kTma && (int)g.c == kTma) {
        tp = (uint32_t)__cvta_generic_to_shared(&tiles[warp][0]);
        if (lane == 0) {
            TileRec *rec = &sm.recs[warp];
            rec->map = reinterpret_cast<unsigned long long>(&tmap);
            rec->x0b = p.x0 * g.c; rec->y0 = p.y0; rec->img = p.img;
        }
        __syncwarp();
    }
    static_assert(kLutSize * 2 == 2048, "get() assumes a 2 KB-aligned, 2 KB LUT per warp");
    const uint32_t recs_s = (uint32_t)__cvta_generic_to_shared(&sm.recs[0]);
#define FLIC_ROWS(C, SG, FM) decode_rows<C, SG, FM, kTma>(rs, lut_s, m2048, dst, (int)p.bwa, active, aligned, lane, fmask, fvals, pf, tp, recs_s)
    if (fmask == 0) {
        switch (g.c) {
            case 1: FLIC_ROWS(1, false, 0); break;
            case 2: FLIC_ROWS(2, false, 0); break;
            case 3: if (sg) FLIC_ROWS(3, true, 0); else FLIC_ROWS(3, false, 0); break;
            default: if (sg) FLIC_ROWS(4, true, 0); else FLIC_ROWS(4, false, 0); break;
        }
    } else if (g.c == 4 && fmask == 8u) {  // opaque (or otherwise constant) alpha: three symbols per pixel
        if (sg) FLIC_ROWS(4, true, 8); else FLIC_ROWS(4, false, 8);
    } else {
        switch (g.c) {
            case 1: FLIC_ROWS(1, false, -1); break;
            case 2: FLIC_ROWS(2, false, -1); break;
            case 3: if (sg) FLIC_ROWS(3, true, -1); else FLIC_ROWS(3, false, -1); break;
            default: if (sg) FLIC_ROWS(4, true, -1); else FLIC_ROWS(4, false, -1); break;
        }
    }
#undef FLIC_ROWS
}

void launch_decode(const uint32_t *d_streams, const unsigned long long *d_offsets, const Geo &g,
                   uint8_t *d_pixels, uint32_t *d_err, const void *tensor_map, cudaStream_t s) {
    uint64_t total = (uint64_t)g.n * g.nb;
    unsigned grid = (unsigned)((total + kDecWarps - 1) / kDecWarps);
    CUtensorMap tm;
    if (tensor_map) memcpy(&tm, tensor_map, sizeof tm); else memset(&tm, 0, sizeof tm);
    if (tensor_map && g.c == 4)
        k_decode<4><<<grid, kDecWarps * 32, 0, s>>>(d_streams, d_offsets, g, d_pixels, d_err, 2048u, kPrefetchChunks, kPrefetchL2Words, tm);
    else if (tensor_map && g.c == 3)
        k_decode<3><<<grid, kDecWarps * 32, 0, s>>>(d_streams, d_offsets, g, d_pixels, d_err, 2048u, kPrefetchChunks, kPrefetchL2Words, tm);
    else
        k_decode<0><<<grid, kDecWarps * 32, 0, s>>>(d_streams, d_offsets, g, d_pixels, d_err, 2048u, kPrefetchChunks, kPrefetchL2Words, tm);
}

}  // namespace flic
