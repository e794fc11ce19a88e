// decode_one.cu — decoder for FLIC_FLAG_ONE_STREAM blocks (sm_100a): ONE contiguous bit-serial Huffman
// stream per block, no per-row sub-streams, nothing in the stream that says where a row (or any symbol but
// the first) starts.  This is the pessimistic case VERDICT r1 item 2 asks to be measured: the only parallelism
// inside a block is what the decoder can recover by itself.
//
// One CTA (256 threads) per block, self-synchronising sub-sequence decode (Weißenberger & Schmidt style):
//   0. the block's stream is copied to shared memory once, coalesced (the only HBM read); warp 0 builds
//      the 2^kL-entry LUT meanwhile;
//   1. the stream is cut into 256 sub-sequences of SW words; every thread decodes its own from its first
//      bit — a guess, wrong more often than not — and records where symbols start in a bit mask;
//   2. thread t restarts from the position where thread t-1's chain really ended and runs until it lands on
//      a bit its guessed chain had marked (Huffman decoding is deterministic: from there on the two chains
//      are one); typically a handful of symbols.  Chains that do not meet inside the sub-sequence push the
//      correction on to the next thread, round by round, until nothing moves;
//   3. symbol counts are popcounts of the masks; a block-wide scan gives every thread its first symbol index;
//   4. every thread decodes its symbols again from its true start into a residual tile (32 rows x row bytes);
//   5. un-prediction as scans: column 0 is a byte-wise prefix down the rows (one warp), a row is a byte-wise
//      prefix along the row (lane-local over four pixels, then a warp scan); rows leave as fully coalesced
//      128-bit stores.
// Cost relative to decode.cu's one-lane-per-row reader: every symbol is looked up a little over twice, and
// the residuals make a round trip through shared memory.
//
// Format: DESIGN.md §FLP0.8 (provisional; not the reference's bitstream).
#include "decode_common.cuh"

namespace flic {

constexpr int kOneThreads = 256;
constexpr int kOneWarps = kOneThreads / 32;

template <int C>
struct OneSmem {
    static constexpr int kMaxWords = kBH * ((kBW * C * kL + 31) / 32);  // longest legal block stream
    static constexpr int kMaskWords = ((kMaxWords + kOneThreads - 1) / kOneThreads) * kOneThreads;
    static constexpr int kTP = kBW * C + 16;  // residual tile row pitch in bytes
    uint32_t sst[kMaxWords + 4];              // the block's stream, then zero words
    union {
        uint32_t mask[kMaskWords];            // steps 1-3: bit b of thread t's words = a symbol starts at bit t*S + b
        uint8_t tile[kBH * kTP];              // steps 4-5: residual bytes, row-major
    } u;
    uint16_t lut[kLutSize];
    LutScratch sc;
    uint32_t E[kOneThreads];                  // where each thread's chain ends = where the next one's starts
    uint32_t wsum[kOneWarps];
    uint32_t vrow[kBH];                       // value of column 0 per row
    uint32_t ok;
};

__device__ __forceinline__ uint32_t peek_bits(const uint32_t *sst, uint32_t pos) {
    const uint32_t wi = pos >> 5;
    return __funnelshift_l(sst[wi + 1], sst[wi], pos);  // the 32 bits from `pos` on, MSB-first
}

// four pixels, C valid low bytes each (higher bytes: junk), from the lane's C packed words
template <int C>
__device__ __forceinline__ void unpack4(const uint32_t *w, uint32_t (&px)[4]) {
    if (C == 4) { px[0] = w[0]; px[1] = w[1 % C]; px[2] = w[2 % C]; px[3] = w[3 % C]; }
    else if (C == 3) {
        px[0] = w[0];
        px[1] = __byte_perm(w[0], w[1 % C], 0x4543);
        px[2] = __byte_perm(w[1 % C], w[2 % C], 0x4432);
        px[3] = w[2 % C] >> 8;
    } else if (C == 2) { px[0] = w[0]; px[1] = w[0] >> 16; px[2] = w[1 % C]; px[3] = w[1 % C] >> 16; }
    else { px[0] = w[0]; px[1] = w[0] >> 8; px[2] = w[0] >> 16; px[3] = w[0] >> 24; }
}

template <int C>
__global__ void __launch_bounds__(kOneThreads) k_decode_one(const uint32_t *__restrict__ streams,
                                                           const unsigned long long *__restrict__ offsets, Geo g,
                                                           uint8_t *__restrict__ pixels, uint32_t *err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    OneSmem<C> &sm = *reinterpret_cast<OneSmem<C> *>(smem_raw);
    constexpr int kTP = OneSmem<C>::kTP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t gb = blockIdx.x;
    const BlockPos p = block_pos(g, gb);

    // ---- stream header, directory entry (every thread reads the same words: uniform control flow)
    const unsigned long long sbeg = offsets[p.img], send = offsets[p.img + 1];
    const uint32_t *sw = streams + (sbeg >> 2);
    const uint64_t swords = (send - sbeg) >> 2;
    const uint64_t fixed = kHdrWords + (uint64_t)g.nb + 1;
    bool ok = (sbeg & 3) == 0 && send >= sbeg && swords >= fixed;
    uint32_t off = 0, end = 0;
    if (ok) {
        const uint32_t pw = sw[6];
        off = sw[kHdrWords + p.b];
        end = sw[kHdrWords + p.b + 1];
        ok = sw[0] == kMagic && sw[1] == (kVersion | (g.c << 16) | ((g.flags & 0xFFu) << 24)) && sw[2] == g.w && sw[3] == g.h &&
             sw[4] == ((uint32_t)kBW | ((uint32_t)kBH << 16)) && sw[5] == g.nb && sw[7] == (uint32_t)kL && fixed + pw <= swords &&
             off <= end && end <= pw && end - off >= (uint32_t)kBlkHdrWords1 &&
             end - off - (uint32_t)kBlkHdrWords1 <= (uint32_t)OneSmem<C>::kMaxWords;
    }
    if (!ok) {
        if (tid == 0) atomicOr(err, kErrFormat);
        return;
    }
    const uint32_t *blk = sw + fixed + off;
    const uint32_t nw = end - off - (uint32_t)kBlkHdrWords1;
    const uint32_t fmask = __ldg(blk + 32), fvals = __ldg(blk + 33);

    // ---- 0. stream -> shared memory; LUT
    for (uint32_t i = tid; i < nw + 4; i += kOneThreads) sm.sst[i] = i < nw ? __ldg(blk + kBlkHdrWords1 + i) : 0u;
    if (warp == 0) {
        const bool lut_ok = build_lut(sm.lut, sm.sc, __ldg(blk + lane), lane);
        if (lane == 0) sm.ok = lut_ok && (fmask >> C) == 0;
    }
    __syncthreads();
    if (!sm.ok) {
        if (tid == 0) atomicOr(err, kErrFormat);
        return;
    }
    const uint32_t CS = (uint32_t)C - __popc(fmask);       // coded symbols per pixel
    const uint32_t nsym = p.bwa * p.bha * CS;
    const bool nobits = (sm.lut[0] & 0xFFu) == 0;           // one symbol with a zero-length code: every LUT entry is it
    const char *lutb = reinterpret_cast<const char *>(sm.lut);

    if (nobits || nsym == 0) {
        // nothing to read: every coded byte is the sole symbol
        const uint32_t fill = (uint32_t)(sm.lut[0] >> 8) * 0x01010101u;
        uint32_t *t32 = reinterpret_cast<uint32_t *>(sm.u.tile);
        for (int i = tid; i < kBH * kTP / 4; i += kOneThreads) t32[i] = fill;
    } else {
        // ---- 1. speculative chains
        const uint32_t nbits = 32u * nw;
        const uint32_t SW = max(1u, (nw + kOneThreads - 1) / kOneThreads), S = 32u * SW;
        const uint32_t start = (uint32_t)tid * S, limit = min(start + S, nbits);
        const bool active = start < nbits;
        uint32_t *mymask = sm.u.mask + (uint32_t)tid * SW;
        uint32_t myend = nbits;
        if (active) {
            uint32_t pos = start, w = 0, cur = 0;
            while (pos < limit) {
                const uint32_t rel = pos - start;
                if ((rel >> 5) != w) {
                    mymask[w] = cur;
                    for (++w; w < (rel >> 5); ++w) mymask[w] = 0;
                    cur = 0;
                }
                cur |= 1u << (rel & 31);
                const uint32_t e = *reinterpret_cast<const uint16_t *>(lutb + ((peek_bits(sm.sst, pos) >> (31 - kL)) & (2 * kLutSize - 2)));
                pos += max(e & 0xFFu, 1u);
            }
            mymask[w] = cur;
            for (++w; w < SW; ++w) mymask[w] = 0;
            myend = pos;
        } else {
            for (uint32_t w = 0; w < SW; ++w) mymask[w] = 0;
        }
        sm.E[tid] = myend;
        __syncthreads();

        // ---- 2. synchronisation rounds
        uint32_t mystart = start;
        for (;;) {
            const uint32_t prev = tid ? sm.E[tid - 1] : 0u;
            bool changed = false;
            if (active && tid && prev != mystart) {
                mystart = prev;
                const uint32_t olde = myend;
                if (prev >= limit) {  // the predecessor's last symbol runs past this whole sub-sequence
                    for (uint32_t w = 0; w < SW; ++w) mymask[w] = 0;
                    myend = prev;
                } else {
                    uint32_t pos = prev, w = (pos - start) >> 5;
                    for (uint32_t i = 0; i < w; ++i) mymask[i] = 0;
                    uint32_t old = mymask[w], cur = 0;
                    bool synced = false;
                    while (pos < limit) {
                        const uint32_t rel = pos - start;
                        if ((rel >> 5) != w) {
                            mymask[w] = cur;
                            for (++w; w < (rel >> 5); ++w) mymask[w] = 0;
                            old = mymask[w];
                            cur = 0;
                        }
                        const uint32_t bit = 1u << (rel & 31);
                        if (old & bit) {  // the guessed chain passed through here: from now on it is the true one
                            mymask[w] = cur | (old & ~(bit - 1u));
                            synced = true;
                            break;
                        }
                        cur |= bit;
                        const uint32_t e = *reinterpret_cast<const uint16_t *>(lutb + ((peek_bits(sm.sst, pos) >> (31 - kL)) & (2 * kLutSize - 2)));
                        pos += max(e & 0xFFu, 1u);
                    }
                    if (!synced) {
                        mymask[w] = cur;
                        for (++w; w < SW; ++w) mymask[w] = 0;
                        myend = pos;
                    }
                }
                changed = myend != olde;
            }
            __syncthreads();  // everyone has read its predecessor's end
            sm.E[tid] = myend;
            if (!__syncthreads_or(changed)) break;
        }

        // ---- 3. symbol counts -> first symbol index of every thread
        uint32_t cnt = 0;
        for (uint32_t w = 0; w < SW; ++w) cnt += __popc(mymask[w]);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        if (lane == 31) sm.wsum[warp] = incl;
        const uint32_t tstart = tid ? sm.E[tid - 1] : 0u;
        __syncthreads();  // all mask reads are done: the tile (same memory) may be written from here on
        uint32_t base = incl - cnt, total = 0;
#pragma unroll
        for (int k = 0; k < kOneWarps; ++k) { const uint32_t s = sm.wsum[k]; base += k < warp ? s : 0u; total += s; }
        if (total < nsym && tid == 0) atomicOr(err, kErrFormat);  // the stream holds fewer symbols than the block has (a few more: its padding)

        // ---- 4. the real decode, into the residual tile
        if (cnt && base < nsym) {
            const uint32_t n = min(cnt, nsym - base);
            const uint32_t pix = base / CS;
            uint32_t k = base - pix * CS, row = pix / p.bwa, x = pix - row * p.bwa;
            uint32_t chmap = 0;  // nibble j: the j-th coded channel
            {
                int j = 0;
#pragma unroll
                for (int ch = 0; ch < C; ++ch)
                    if (!((fmask >> ch) & 1u)) chmap |= (uint32_t)ch << (4 * j++);
            }
            uint8_t *tp = sm.u.tile + row * kTP + x * C;
            uint32_t pos = tstart;
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t e = *reinterpret_cast<const uint16_t *>(lutb + ((peek_bits(sm.sst, pos) >> (31 - kL)) & (2 * kLutSize - 2)));
                pos += max(e & 0xFFu, 1u);
                tp[(chmap >> (4 * k)) & 15u] = (uint8_t)(e >> 8);
                if (++k == CS) {
                    k = 0; tp += C;
                    if (++x == p.bwa) { x = 0; ++row; tp = sm.u.tile + row * kTP; }
                }
            }
        }
    }
    __syncthreads();

    // ---- 5. un-prediction and stores
    const uint32_t fb = ((fmask & 1u) ? 0xFFu : 0u) | ((fmask & 2u) ? 0xFF00u : 0u) | ((fmask & 4u) ? 0xFF0000u : 0u) |
                        ((fmask & 8u) ? 0xFF000000u : 0u);
    const uint32_t keep = ~fb, fl = fvals & fb;
    const bool sg = (g.flags & FLIC_FLAG_SUBGREEN) != 0 && C >= 3;
    if (warp == 0) {  // column 0: pixel (r, 0) = sum of the first residuals of rows 0..r
        uint32_t fp = 0;
        if (lane < (int)p.bha) {
            const uint8_t *t = sm.u.tile + lane * kTP;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) fp |= (uint32_t)t[ch] << (8 * ch);
        }
        fp &= keep;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, fp, d);
            if (lane >= d) fp = __vadd4(fp, t);
        }
        sm.vrow[lane] = fp;
    }
    __syncthreads();
    const int npx = max(0, min(4, (int)p.bwa - 4 * lane));
    for (int r = warp; r < (int)p.bha; r += kOneWarps) {
        const uint32_t above = r ? sm.vrow[r - 1] : 0u;
        uint32_t w[C], px[4];
        const uint32_t *t = reinterpret_cast<const uint32_t *>(sm.u.tile + r * kTP + 4 * C * lane);
        if (C == 4) {
            const uint4 q = *reinterpret_cast<const uint4 *>(t);
            w[0] = q.x; w[1 % C] = q.y; w[2 % C] = q.z; w[3 % C] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < C; ++j) w[j] = t[j];
        }
        unpack4<C>(w, px);
        px[0] &= keep;  // flat channels carry no residuals: whatever the tile holds there stays out of the sums
        px[1] = __vadd4(px[0], px[1] & keep);
        px[2] = __vadd4(px[1], px[2] & keep);
        px[3] = __vadd4(px[2], px[3] & keep);
        // the lane's last real pixel (lanes past the block's right edge add nothing)
        uint32_t run = npx >= 4 ? px[3] : (npx == 3 ? px[2] : (npx == 2 ? px[1] : (npx == 1 ? px[0] : 0u)));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, run, d);
            if (lane >= d) run = __vadd4(run, u);
        }
        uint32_t before = __shfl_up_sync(0xFFFFFFFFu, run, 1);
        if (lane == 0) before = 0;
        before = __vadd4(before, above);
        uint32_t o[C];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t v = (__vadd4(px[i], before) & keep) | fl;
            if (sg) v = addgreen4(v);
            px[i] = v;
        }
        pack4<C>(px[0], px[1], px[2], px[3], o);
        uint8_t *dst = pixels + (uint64_t)p.img * g.img_stride + (uint64_t)(p.y0 + r) * g.pitch + (uint64_t)p.x0 * C + 4 * C * lane;
        if (npx == 4 && g.aligned16) {
            if (C == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1 % C], o[2 % C], o[3 % C]);
            else if (C == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1 % C]);
            else {
#pragma unroll
                for (int j = 0; j < C; ++j) reinterpret_cast<uint32_t *>(dst)[j] = o[j];
            }
        } else {
            for (int i = 0; i < npx * C; ++i) dst[i] = (uint8_t)(o[i >> 2] >> (8 * (i & 3)));
        }
    }
}

void launch_decode_one(const uint32_t *d_streams, const unsigned long long *d_offsets, const Geo &g, uint8_t *d_pixels,
                       uint32_t *d_err, cudaStream_t s) {
    const uint64_t total = (uint64_t)g.n * g.nb;
#define FLIC_ONE(C)                                                                                                  \
    do {                                                                                                             \
        static bool attr = false;                                                                                    \
        if (!attr) {                                                                                                 \
            cudaFuncSetAttribute(k_decode_one<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OneSmem<C>)); \
            attr = true;                                                                                             \
        }                                                                                                            \
        k_decode_one<C><<<(unsigned)total, kOneThreads, sizeof(OneSmem<C>), s>>>(d_streams, d_offsets, g, d_pixels, d_err); \
    } while (0)
    switch (g.c) {
        case 1: FLIC_ONE(1); break;
        case 2: FLIC_ONE(2); break;
        case 3: FLIC_ONE(3); break;
        default: FLIC_ONE(4); break;
    }
#undef FLIC_ONE
}

}  // namespace flic
