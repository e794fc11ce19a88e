// encode.cu — the four encode kernels of the FLP0 engine (sm_100a).
//
//   k_histograms : one CTA per block; residuals computed in registers from
//                  128-bit coalesced row loads, warp-privatised shared-memory
//                  histograms, 512 B of u16 counts written per block.
//   k_tables     : one WARP per block; bitonic sort of (count,symbol) keys,
//                  two-queue Huffman merge, depth census, Kraft repair,
//                  canonical code assignment by packed-counter warp scan.
//   k_pack       : one CTA per block; residuals recomputed in registers (the
//                  tile never touches shared memory), per-lane code
//                  concatenation, warp scan of bit lengths, OR-scatter into a
//                  shared staging tile, decoupled look-back over block sizes,
//                  coalesced copy-out straight into the final stream position.
//   k_finalize   : headers, rebased u32 directories and the n+1 stream offsets.
//
// Format: DESIGN.md §FLP0 (provisional; not the reference's bitstream —
// LICENSING.md).  Byte-exact CPU model: oracle/flp0_oracle.c (tests only).
#include "common.cuh"

namespace flic {

// byte j of w, times four (a u32 table offset), in two instructions
__device__ __forceinline__ uint32_t byte_x4(uint32_t w, int j) {
    return j == 0 ? (w << 2) & 0x3FCu : (w >> (8 * j - 2)) & 0x3FCu;
}

// ---------------------------------------------------------------- k_histograms
constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;

__global__ void __launch_bounds__(kEncThreads) k_histograms(const uint8_t *__restrict__ pixels, Geo g,
                                                            uint16_t *__restrict__ hist) {
    __shared__ uint32_t sh[kEncWarps][256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t gb = blockIdx.x;
    const BlockPos p = block_pos(g, gb);

    for (int i = tid; i < kEncWarps * 256; i += kEncThreads) (&sh[0][0])[i] = 0;

    uint4 res[kBH / kEncWarps];
    int nv[kBH / kEncWarps];
#pragma unroll
    for (int q = 0; q < kBH / kEncWarps; ++q) {
        int r = warp + kEncWarps * q;
        nv[q] = 0;
        res[q] = make_uint4(0, 0, 0, 0);
        if (r < (int)p.bha) res[q] = row_residuals(pixels, g, p, r, lane, &nv[q]);  // warp-uniform branch
    }
    __syncthreads();
    char *my = reinterpret_cast<char *>(sh[warp]);
#pragma unroll
    for (int q = 0; q < kBH / kEncWarps; ++q) {
        const uint32_t w[4] = {res[q].x, res[q].y, res[q].z, res[q].w};
        if (nv[q] == 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(reinterpret_cast<uint32_t *>(my + byte_x4(w[j >> 2], j & 3)), 1u);
        } else if (nv[q] > 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < nv[q]) atomicAdd(reinterpret_cast<uint32_t *>(my + byte_x4(w[j >> 2], j & 3)), 1u);
        }
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kEncWarps; ++k) s += sh[k][tid];
    hist[gb * 256 + tid] = (uint16_t)s;
}

void launch_histograms(const uint8_t *d_pixels, const Geo &g, uint16_t *d_hist, cudaStream_t s) {
    uint64_t total = (uint64_t)g.n * g.nb;
    k_histograms<<<(unsigned)total, kEncThreads, 0, s>>>(d_pixels, g, d_hist);
}

// -------------------------------------------------------------------- k_tables
constexpr int kTabWarps = 8;

struct TableScratch {
    uint32_t key[256];  // (count << 8) | symbol, sorted ascending; unused symbols = 0xFFFFFFFF
    uint32_t iw[256];   // internal-node weights in creation order
    uint8_t lpar[256];  // parent (internal index) of sorted leaf i
    uint8_t ipar[256];  // parent of internal node k
    uint8_t idep[256];  // depth of internal node k (root = 0)
    uint8_t len[256];   // code length per SYMBOL
    uint32_t num[16];   // leaves per length
    uint32_t cum[16];   // ranks < cum[l] get a length >= l
    uint32_t next[16];  // first canonical code of each length
};

// 12 nine-bit counters (lengths 1..12) packed into two u64
__device__ __forceinline__ void cnt_add(uint64_t &a, uint64_t &b, uint32_t l) {
    if (l >= 1 && l <= 6) a += 1ull << (9 * (l - 1));
    else if (l >= 7 && l <= 12) b += 1ull << (9 * (l - 7));
}
__device__ __forceinline__ uint32_t cnt_get(uint64_t a, uint64_t b, uint32_t l) {
    return (uint32_t)((l <= 6 ? a >> (9 * (l - 1)) : b >> (9 * (l - 7))) & 511u);
}
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d) {
    uint32_t lo = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)v, d);
    uint32_t hi = __shfl_up_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), d);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(kTabWarps * 32) k_tables(const uint16_t *__restrict__ hist, uint64_t nblocks,
                                                           uint16_t *__restrict__ table) {
    __shared__ TableScratch S[kTabWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t gb = (uint64_t)blockIdx.x * kTabWarps + warp;
    if (gb >= nblocks) return;  // whole warp leaves; no block-level sync below
    TableScratch &s = S[warp];

    // FLP0 §3.1: keys of the lane's 8 symbols
    const uint4 hv = *reinterpret_cast<const uint4 *>(hist + gb * 256 + 8 * lane);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
    uint32_t nact = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t f = (hw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu, sym = 8 * lane + k;
        s.key[sym] = f ? ((f << 8) | sym) : 0xFFFFFFFFu;
        s.len[sym] = 0;
        nact += f != 0;
    }
    if (lane < 16) s.num[lane] = 0;
    const int n = (int)warp_sum(nact);
    __syncwarp();

    if (n >= 2) {
        // bitonic sort, ascending by (count, symbol)
        for (int k = 2; k <= 256; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int t = lane; t < 128; t += 32) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), q = i | j;
                    uint32_t a = s.key[i], b = s.key[q];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { s.key[i] = b; s.key[q] = a; }
                }
                __syncwarp();
            }
        }
        // FLP0 §3.2: two-queue merge; a leaf wins a tie against an internal node
        if (lane == 0) {
            int li = 0, ii = 0;
            for (int k = 0; k < n - 1; ++k) {
                uint32_t wsum = 0;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    bool leaf = li < n && (ii >= k || (s.key[li] >> 8) <= s.iw[ii]);
                    if (leaf) { wsum += s.key[li] >> 8; s.lpar[li++] = (uint8_t)k; }
                    else { wsum += s.iw[ii]; s.ipar[ii++] = (uint8_t)k; }
                }
                s.iw[k] = wsum;
            }
            s.idep[n - 2] = 0;
            for (int k = n - 3; k >= 0; --k) s.idep[k] = (uint8_t)(s.idep[s.ipar[k]] + 1);
        }
        __syncwarp();
        // FLP0 §3.3: leaves per depth, depths beyond L folded into L
        for (int i = lane; i < n; i += 32) {
            uint32_t d = (uint32_t)s.idep[s.lpar[i]] + 1u;
            atomicAdd(&s.num[min(d, (uint32_t)kL)], 1u);
        }
        __syncwarp();
        // FLP0 §3.4: Kraft repair, then rank boundaries and first codes
        if (lane == 0) {
            uint32_t total = 0;
            for (int l = kL; l >= 1; --l) total += s.num[l] << (kL - l);
            while (total > (1u << kL)) {
                s.num[kL]--;
                for (int l = kL - 1; l >= 1; --l)
                    if (s.num[l]) { s.num[l]--; s.num[l + 1] += 2; break; }
                total--;
            }
            uint32_t c = 0;
            for (int l = kL; l >= 1; --l) { c += s.num[l]; s.cum[l] = c; }
            s.next[1] = 0;
            for (int l = 2; l <= kL; ++l) s.next[l] = (s.next[l - 1] + s.num[l - 1]) << 1;
        }
        __syncwarp();
        // FLP0 §3.5: lengths by sorted rank (rarest first -> longest)
        for (int i = lane; i < n; i += 32) {
            uint32_t l = 0;
#pragma unroll
            for (int t = 1; t <= kL; ++t)
                if ((uint32_t)i < s.cum[t]) l = t;
            s.len[s.key[i] & 0xFFu] = (uint8_t)l;
        }
        __syncwarp();
    } else if (n == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (s.key[8 * lane + k] != 0xFFFFFFFFu) s.len[8 * lane + k] = (uint8_t)kLenSole;
        __syncwarp();
    }

    // FLP0 §4: canonical codes in (length, symbol) order; the lane owns symbols 8*lane..8*lane+7
    uint32_t l8[8];
    uint64_t ca = 0, cb = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { l8[k] = s.len[8 * lane + k]; cnt_add(ca, cb, l8[k]); }
    uint64_t ia = ca, ib = cb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t ta = shfl_up64(ia, d), tb = shfl_up64(ib, d);
        if (lane >= d) { ia += ta; ib += tb; }
    }
    uint64_t ea = ia - ca, eb = ib - cb;  // symbols of each length in lower lanes
    uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t l = l8[k], e = 0;
        if (l >= 1 && l <= (uint32_t)kL) {
            e = (l << 12) | (s.next[l] + cnt_get(ea, eb, l));
            cnt_add(ea, eb, l);
        } else if (l == kLenSole) {
            e = kLenSole << 12;
        }
        out[k >> 1] |= e << (16 * (k & 1));
    }
    *reinterpret_cast<uint4 *>(table + gb * 256 + 8 * lane) = make_uint4(out[0], out[1], out[2], out[3]);
}

void launch_tables(const uint16_t *d_hist, uint64_t nblocks, uint16_t *d_table, cudaStream_t s) {
    unsigned grid = (unsigned)((nblocks + kTabWarps - 1) / kTabWarps);
    k_tables<<<grid, kTabWarps * 32, 0, s>>>(d_hist, nblocks, d_table);
}

// ---------------------------------------------------------------------- k_pack
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValMask = (1ull << 62) - 1;
constexpr uint32_t kSpinLimit = 1u << 24;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}

// Single-pass chained scan (decoupled look-back) over block payload sizes, run by warp 0.
// Returns the exclusive prefix (words) of block gb; blocks are claimed in ticket order, so
// every predecessor is already running and the wait is bounded.
__device__ unsigned long long lookback(unsigned long long *status, uint64_t gb, uint32_t size, int lane,
                                       uint32_t *err) {
    if (lane == 0) st_status(status + gb, kFlagAgg | size);
    unsigned long long excl = 0;
    long long idx = (long long)gb - 1;
    while (idx >= 0) {
        long long my = idx - lane;
        unsigned long long v;
        uint32_t spins = 0;
        bool pending;
        do {
            v = my >= 0 ? ld_status(status + my) : kFlagPrefix;
            pending = __any_sync(0xFFFFFFFFu, (v >> 62) == 0);
        } while (pending && ++spins < kSpinLimit);
        if (pending) {
            if (lane == 0) atomicOr(err, kErrWatchdog);
            break;
        }
        uint32_t pm = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
        unsigned long long val = v & kValMask;
        if (pm) {
            int first = __ffs(pm) - 1;
            if (lane > first) val = 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            uint32_t lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)val, d);
            uint32_t hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(val >> 32), d);
            val += ((unsigned long long)hi << 32) | lo;
        }
        excl += val;
        if (pm) break;
        idx -= 32;
    }
    if (lane == 0) st_status(status + gb, kFlagPrefix | ((excl + size) & kValMask));
    return excl;
}

constexpr int kStagePitch = kRowWordsMax + 1;  // odd pitch: the interleaving copy-out reads a column conflict-free

// Looks up the lane's 16 symbols and merges them pairwise: pk[i] = (bits << 24) | code bits of
// symbols 2i,2i+1 (at most 22 bits).  Returns the lane's total bit count.
template <bool kFull>
__device__ __forceinline__ uint32_t gather_pairs(const uint32_t (&w)[4], int nv, const uint32_t *tab, uint32_t (&pk)[8]) {
    const char *t = reinterpret_cast<const char *>(tab);
    uint32_t nbits = 0;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        uint32_t e0 = *reinterpret_cast<const uint32_t *>(t + byte_x4(w[j >> 2], j & 3));
        uint32_t e1 = *reinterpret_cast<const uint32_t *>(t + byte_x4(w[j >> 2], (j & 3) + 1));
        if (!kFull) {
            if (j >= nv) e0 = 0;
            if (j + 1 >= nv) e1 = 0;
        }
        const uint32_t l1 = e1 >> 16, l = (e0 >> 16) + l1;
        pk[j >> 1] = (((e0 & 0xFFFFu) << l1) | (e1 & 0xFFFFu)) | (l << 24);
        nbits += l;
    }
    return nbits;
}

__global__ void __launch_bounds__(kEncThreads) k_pack(const uint8_t *__restrict__ pixels, Geo g,
                                                      const uint16_t *__restrict__ table,
                                                      uint32_t *__restrict__ streams, uint64_t capacity_words,
                                                      unsigned long long *status, unsigned long long *dirE,
                                                      uint32_t *err) {
    __shared__ __align__(16) uint32_t stage[kBH * kStagePitch];
    __shared__ uint32_t tab[256];  // (len << 16) | code; len 0 for a sole symbol
    __shared__ uint8_t nib[256];
    __shared__ uint32_t rwc[kBH], rowoff[kBH];
    __shared__ unsigned long long s_gb, s_base;
    __shared__ uint32_t s_minw;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t total_blocks = (uint64_t)g.n * g.nb;

    // blocks are claimed in launch order so that look-back predecessors are always resident or done
    if (tid == 0) s_gb = atomicAdd(status + total_blocks, 1ull);
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *s4 = reinterpret_cast<uint4 *>(stage);
        for (int i = tid; i < kBH * kStagePitch / 4; i += kEncThreads) s4[i] = z;
    }
    __syncthreads();
    const uint64_t gb = s_gb;
    const BlockPos p = block_pos(g, gb);

    {
        uint32_t e = table[gb * 256 + tid], l = e >> 12;
        nib[tid] = (uint8_t)l;
        tab[tid] = l == kLenSole ? 0u : ((l << 16) | (e & 0xFFFu));
    }
    uint4 res[kBH / kEncWarps];
    int nv[kBH / kEncWarps];
#pragma unroll
    for (int q = 0; q < kBH / kEncWarps; ++q) {
        int r = warp + kEncWarps * q;
        nv[q] = 0;
        res[q] = make_uint4(0, 0, 0, 0);
        if (r < (int)p.bha) res[q] = row_residuals(pixels, g, p, r, lane, &nv[q]);
    }
    __syncthreads();

    // FLP0 §5: one warp per row.  Each lane merges its 16 codes pairwise, a warp scan of bit
    // counts places them, and 32-bit words are OR-scattered into the zeroed staging row.
#pragma unroll
    for (int q = 0; q < kBH / kEncWarps; ++q) {
        const int r = warp + kEncWarps * q;
        const uint32_t w[4] = {res[q].x, res[q].y, res[q].z, res[q].w};
        uint32_t pk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nbits = 0;
        if (nv[q] == 16) nbits = gather_pairs<true>(w, 16, tab, pk);
        else if (nv[q] > 0) nbits = gather_pairs<false>(w, nv[q], tab, pk);
        const uint32_t incl = warp_incl_scan(nbits, lane);
        if (nbits) {
            const uint32_t o = incl - nbits;
            uint32_t *dst = &stage[r * kStagePitch + (o >> 5)];
            unsigned long long acc = 0;
            uint32_t na = o & 31u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc = (acc << (pk[i] >> 24)) | (pk[i] & 0xFFFFFFu);
                na += pk[i] >> 24;
                if (na >= 32u) {
                    na -= 32u;
                    atomicOr(dst, (uint32_t)(acc >> na));
                    ++dst;
                }
            }
            if (na) atomicOr(dst, (uint32_t)(acc << (32u - na)));
        }
        if (lane == 31) rwc[r] = (incl + 31u) >> 5;
    }
    __syncthreads();

    if (warp == 0) {
        const uint32_t wcount = rwc[lane];
        const uint32_t incl = warp_incl_scan(wcount, lane);
        rowoff[lane] = incl - wcount;
        uint32_t mn = lane < (int)p.bha ? wcount : 0xFFFFFFFFu;  // FLP0 §6: interleave depth
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
        const uint32_t size = kBlkHdrWords + __shfl_sync(0xFFFFFFFFu, incl, 31);
        const unsigned long long excl = lookback(status, gb, size, lane, err);
        if (lane == 0) {
            s_minw = mn;
            dirE[gb] = excl;
            if (gb == total_blocks - 1) dirE[total_blocks] = excl + size;
            unsigned long long base = (unsigned long long)(p.img + 1) * (kHdrWords + g.nb + 1) + excl;
            if (base + size > capacity_words) {
                atomicOr(err, kErrCapacity);
                base = ~0ull;
            }
            s_base = base;
        }
    }
    __syncthreads();
    if (s_base == ~0ull) return;
    uint32_t *out = streams + s_base;
    if (warp == 0) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v |= (uint32_t)nib[8 * lane + k] << (4 * k);
        out[lane] = v;
    } else if (warp == 1 && lane < kBH / 2) {
        out[32 + lane] = rwc[2 * lane] | (rwc[2 * lane + 1] << 16);
    }
    // interleaved region: word k of row r lands at k*bha + r (coalesced stores, conflict-free column reads)
    const uint32_t minw = s_minw, bha = p.bha, inter = minw * bha;
    out += kBlkHdrWords;
    if (bha == (uint32_t)kBH) {
        for (uint32_t i = tid; i < inter; i += kEncThreads) out[i] = stage[(i & 31u) * kStagePitch + (i >> 5)];
    } else {
        for (uint32_t i = tid; i < inter; i += kEncThreads) {
            uint32_t k = i / bha, r = i - k * bha;
            out[i] = stage[r * kStagePitch + k];
        }
    }
    // tails, row by row
    for (uint32_t r = warp; r < bha; r += kEncWarps) {
        const uint32_t cnt = rwc[r] - minw;
        uint32_t *o = out + inter + (rowoff[r] - r * minw);
        const uint32_t *src = &stage[r * kStagePitch + minw];
        for (uint32_t i = lane; i < cnt; i += 32) o[i] = src[i];
    }
}

void launch_pack(const uint8_t *d_pixels, const Geo &g, const uint16_t *d_table, uint32_t *d_streams,
                 uint64_t capacity_words, unsigned long long *d_status, unsigned long long *d_dirE,
                 uint32_t *d_err, cudaStream_t s) {
    uint64_t total = (uint64_t)g.n * g.nb;  // caller has zeroed d_status[0..total] on this stream
    k_pack<<<(unsigned)total, kEncThreads, 0, s>>>(d_pixels, g, d_table, d_streams, capacity_words, d_status,
                                                  d_dirE, d_err);
}

// ------------------------------------------------------------------ k_finalize
__global__ void __launch_bounds__(256) k_finalize(Geo g, const unsigned long long *__restrict__ dirE,
                                                  uint32_t *__restrict__ streams, uint64_t capacity_words,
                                                  unsigned long long *__restrict__ offsets) {
    const uint64_t per = kHdrWords + (uint64_t)g.nb + 1;  // header + directory words per image
    const uint64_t items = (uint64_t)g.n * per;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < items + g.n + 1;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (i >= items) {  // stream offsets in bytes
            uint64_t img = i - items;
            offsets[img] = 4ull * (img * per + dirE[img * g.nb]);
            continue;
        }
        uint64_t img = i / per, k = i - img * per;
        unsigned long long first = dirE[img * g.nb];
        uint64_t pos = img * per + first + k;
        if (pos >= capacity_words) continue;  // k_pack already raised kErrCapacity
        uint32_t v;
        if (k >= kHdrWords) {
            v = (uint32_t)(dirE[img * g.nb + (k - kHdrWords)] - first);
        } else {
            switch (k) {
                case 0: v = kMagic; break;
                case 1: v = 2u | (g.c << 16) | ((g.flags & 0xFFu) << 24); break;
                case 2: v = g.w; break;
                case 3: v = g.h; break;
                case 4: v = (uint32_t)kBW | ((uint32_t)kBH << 16); break;
                case 5: v = g.nb; break;
                case 6: v = (uint32_t)(dirE[(img + 1) * g.nb] - first); break;
                default: v = (uint32_t)kL; break;
            }
        }
        streams[pos] = v;
    }
}

void launch_finalize(const Geo &g, const unsigned long long *d_dirE, uint32_t *d_streams,
                     uint64_t capacity_words, unsigned long long *d_offsets, uint32_t *, cudaStream_t s) {
    uint64_t items = (uint64_t)g.n * (kHdrWords + (uint64_t)g.nb + 1) + g.n + 1;
    uint64_t want = (items + 255) / 256;
    unsigned grid = (unsigned)(want < 148ull * 8 ? want : 148ull * 8);
    k_finalize<<<grid, 256, 0, s>>>(g, d_dirE, d_streams, capacity_words, d_offsets);
}

}  // namespace flic
