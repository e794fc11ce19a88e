// encode.cu — the encode kernels of the FLP0 engine (sm_100a): a staged pipeline of five kernels and a fused
// single-pass kernel that produce identical bytes in every layout (DESIGN.md §4.1, §4.2, §5.2 for when each is used and why).
//
//  staged (large batches: every stage at full occupancy, the serial Huffman merges of eight blocks share a warp)
//   k_histograms : one CTA per block, grid (nbx, nby, n); the block's pixels arrive as ONE TMA tile load
//                  (cp.async.bulk.tensor + mbarrier; direct 128-bit loads when the batch is not 16-byte aligned),
//                  residuals in registers, written to the residual plane with ONE bulk store, 2-4 shared sub-histograms
//                  whose increments address with one LOP3, flat-channel detection (FLP0 §2b), 512 B of u16 counts per block.
//   k_tables     : one WARP per 8 blocks, four warps per CTA; bitonic sort of (count,symbol) keys, two-queue Huffman
//                  merge (one block per lane), depth census, Kraft repair, lengths by rank, canonical code assignment by
//                  packed-counter warp scan, and the block's total of count x code length.
//   k_slots      : exclusive prefix sum of block sizes where they follow from the histograms (slots: FLP0 §7; ONE_STREAM:
//                  exact): every block's output position is known before anything is packed.
//   k_pack       : one CTA per block; reads the residual plane, merges symbols into code groups (three per group and two
//                  groups per placement for opaque-alpha RGBA), warp scan of bit lengths, RED.OR of every group into a
//                  zeroed shared staging tile at its end bit position; copy-out by layout: interleaved rows into the
//                  block's slot / the rows concatenated bit-exactly (ONE_STREAM) / ticket order + decoupled look-back
//                  over the packed sizes (EXACT); optionally into a larger image's spliced stream on another GPU
//                  (block-row split through peer memory).
//   k_finalize   : headers, rebased u32 directories and the n+1 stream offsets.
//
//  fused (small jobs)
//   k_encode     : persistent CTAs; per block: rows -> residuals parked in shared memory -> histogram -> code table
//                  built by the whole CTA (cta_table) -> rows packed over the same tile -> position by a decoupled
//                  look-back over the predecessors' sizes -> copy-out.  Pixels are read once; nothing but the
//                  stream is written.
//
// Format: DESIGN.md §FLP0 (provisional; not the reference's bitstream —
// LICENSING.md).  Byte-exact CPU model: oracle/flp0_oracle.c (tests only).
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include <cuda.h>  // CUtensorMap (type only; the driver entry point is resolved in api.cu)

#include "common.cuh"

namespace flic {

// ---------------------------------------------------------------- k_histograms
constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;

// Sub-histograms per CTA: warp w counts into sub-histogram w % kSubHist.  What bounds the kernel is the LSU data pipe
// (ncu: l1tex__data_pipe_lsu_wavefronts 73 % of peak), and eight private copies cost 64 wavefronts to clear and 64 to sum,
// an eighth of the CTA's shared-memory traffic; shared atomics from different warps are separate instructions either way.
// Measured (k_histograms ms, 8 / 4 / 2 copies): 4K RGBA 0.691 / 0.689 / 0.683, noise 1.129 / 1.127 / 1.112, 1080p RGB
// 0.334 / 0.331 / 0.339.
template <int C> struct SubHist { static constexpr int k = C == 3 ? 4 : 2; };

// ++hist[byte j of w], `base_s` = shared-window address of the warp's 1 KB-aligned sub-histogram: the shift, ONE
// LOP3 for (x & 0x3FC) | base and the ATOMS.POPC.INC — a generic pointer + offset cost one more add per symbol.
__device__ __forceinline__ void hist_inc(uint32_t base_s, uint32_t w, int j) {
    const uint32_t x = j == 0 ? (w << 2) : (w >> (8 * j - 2));
    const uint32_t a = (x & 0x3FCu) | base_s;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
}

// resid (optional): the "residual plane" — per block a tile [32 rows][32 lanes][C words] of residual
// bytes in the lane order above — so that k_pack does not recompute prediction (the kernels are
// ALU-bound, HBM has headroom: one extra N-byte write buys ~20 % of k_pack's instructions).
// Per word j of a lane's C words: which of its four bytes belong to a channel named in `mask`
// (byte b of word j carries channel (4j + b) mod C).
template <int C>
__device__ __forceinline__ uint32_t word_channel_bits(uint32_t mask, int j) {
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) r |= ((mask >> ((4 * j + b) % C)) & 1u) << b;
    return r;
}

// flat (optional): per block {mask of flat channels, their values} (FLP0 §2b).  A channel is flat when all
// its residuals are zero except the block's first pixel's, i.e. the OR of the others is zero; the histogram
// then leaves the channel out (bwa*bha - 1 zeros and the first pixel's value are taken back off).
//
// kTma: the block's pixels (32 rows x 128*C bytes) arrive as ONE TMA tile load (cp.async.bulk.tensor, 3-D map
// {row words, rows, images} of u32 elements, completion on an mbarrier) issued by one thread while the others
// clear the sub-histograms; lanes then take their 4*C bytes from shared memory.  Tiles that hang over the image's
// right or bottom edge are zero-filled by the TMA unit, which is exactly what the direct path's edge handling
// produces, and the pixel above a row's first pixel is in the tile too.  Needs a 16-byte-aligned batch.
template <int C, bool SG, bool kTma>
__global__ void __launch_bounds__(kEncThreads, 6) k_histograms(const uint8_t *__restrict__ pixels, Geo g,
                                                            uint16_t *__restrict__ hist, uint32_t *__restrict__ resid,
                                                            uint2 *__restrict__ flat, const __grid_constant__ CUtensorMap tmap,
                                                            int grid3) {
    constexpr int kSubHist = SubHist<C>::k;
    static_assert(kSubHist >= 1 && kSubHist <= kEncWarps && (kSubHist & (kSubHist - 1)) == 0, "power of two");
    __shared__ __align__(1024) uint32_t sh[kSubHist][256];  // 1 KB-aligned rows: hist_inc() ORs the offset into the base
    __shared__ __align__(128) uint32_t ptile[kTma ? kBH * 32 * C : 4];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t s_or[C], s_first;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t gb;
    const BlockPos p = block_pos_cta(g, grid3 != 0, gb);
    if (kTma && tid == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar), dst = (uint32_t)__cvta_generic_to_shared(ptile);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(kBH * 128 * C)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"((int)(p.x0 * C / 4)), "r"((int)p.y0), "r"((int)p.img), "r"(bar)
                     : "memory");
    }

    {
        uint4 *z = reinterpret_cast<uint4 *>(&sh[0][0]);
        if (kSubHist * 256 / 4 >= kEncThreads) {
#pragma unroll
            for (int i = 0; i < kSubHist * 256 / 4 / kEncThreads; ++i) z[tid + i * kEncThreads] = make_uint4(0, 0, 0, 0);
        } else if (tid < kSubHist * 256 / 4) z[tid] = make_uint4(0, 0, 0, 0);
    }
    if (tid < C) s_or[tid] = 0;
    if (kTma) {
        __syncthreads();  // the barrier's initialisation is visible to every thread before anyone waits on it
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
            "@P1 bra DONE_%=;\n\t"
            "bra WAIT_%=;\n\t"
            "DONE_%=:\n\t}" ::"r"(bar) : "memory");
    }

    uint32_t res[kBH / kEncWarps][C];
    uint32_t orw[C];
    int nv[kBH / kEncWarps];
    const bool fast = g.aligned16 != 0;
    const uint8_t *row = pixels + (uint64_t)p.img * g.img_stride + (uint64_t)(p.y0 + warp) * g.pitch + (uint64_t)p.x0 * C;
#pragma unroll
    for (int j = 0; j < C; ++j) orw[j] = 0;
#pragma unroll
    for (int q = 0; q < kBH / kEncWarps; ++q) {
        const int r = warp + kEncWarps * q;
        nv[q] = 0;
#pragma unroll
        for (int j = 0; j < C; ++j) res[q][j] = 0;
        if (r < (int)p.bha) {  // warp-uniform
            uint32_t v[C];
            uint32_t up = 0u;
            if (kTma) {
                const uint32_t *t = ptile + (r * 32 + lane) * C;
                if (C == 4) {
                    const uint4 x = *reinterpret_cast<const uint4 *>(t);
                    v[0] = x.x; v[1 % C] = x.y; v[2 % C] = x.z; v[3 % C] = x.w;
                } else {
#pragma unroll
                    for (int j = 0; j < C; ++j) v[j] = t[j];
                }
                nv[q] = C * max(0, min(4, (int)p.bwa - 4 * lane));
                if (lane == 0 && r > 0) {  // the pixel above this row's first pixel, transformed like the others
                    up = ptile[(r - 1) * 32 * C];
                    if (C < 4) up &= (1u << (8 * (C & 3))) - 1u;
                    if (SG && C >= 3) { const uint32_t gg = (up >> 8) & 0xFFu; up = __vsub4(up, gg | (gg << 16)); }
                }
            } else {
                load_lane_pixels<C>(row, lane, (int)p.bwa, fast, v, &nv[q]);
                if (lane == 0 && r > 0) up = up_pixel<C, SG>(row, g.pitch, fast);
            }
            lane_residuals<C, SG>(v, up, lane, res[q]);
#pragma unroll
            for (int j = 0; j < C; ++j) {
                uint32_t x = res[q][j];
                if (q == 0 && j == 0 && tid == 0) {  // the block's first pixel: its residual is its value
                    constexpr uint32_t fm = C == 4 ? 0xFFFFFFFFu : ((1u << (8 * (C & 3))) - 1u);
                    s_first = x & fm;
                    x &= ~fm;
                }
                orw[j] |= x;
            }
        }
        if (resid && !kTma) {
            uint32_t *t = resid + ((gb * kBH + r) * 32 + lane) * C;
            if (C == 4) *reinterpret_cast<uint4 *>(t) = make_uint4(res[q][0], res[q][1 % C], res[q][2 % C], res[q][3 % C]);
            else if (C == 2) *reinterpret_cast<uint2 *>(t) = make_uint2(res[q][0], res[q][1 % C]);
            else {
#pragma unroll
                for (int j = 0; j < C; ++j) t[j] = res[q][j];
            }
        }
        row += kEncWarps * g.pitch;
    }
    __syncthreads();
    if (kTma && resid) {
        // every read of the pixel tile is done: the residuals take its place (same layout as the residual plane's
        // tile of this block) and leave as ONE bulk store instead of four 128-bit stores per thread
#pragma unroll
        for (int q = 0; q < kBH / kEncWarps; ++q) {
            uint32_t *t = ptile + ((warp + kEncWarps * q) * 32 + lane) * C;
            if (C == 4) *reinterpret_cast<uint4 *>(t) = make_uint4(res[q][0], res[q][1 % C], res[q][2 % C], res[q][3 % C]);
            else {
#pragma unroll
                for (int j = 0; j < C; ++j) t[j] = res[q][j];
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
    }
    {   // bytes past the lane's last real pixel hold garbage differences: keep them out of the OR
        const int nvl = C * max(0, min(4, (int)p.bwa - 4 * lane));
        constexpr int NA = C == 4 ? 1 : C;  // RGBA: byte b is channel b in every word, one accumulator does
        uint32_t acc[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) acc[j] = 0;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const int nb = min(4, max(0, nvl - 4 * j));
            const uint32_t vm = nb == 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
            acc[C == 4 ? 0 : j] |= orw[j] & vm;
        }
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const uint32_t o = __reduce_or_sync(0xFFFFFFFFu, acc[j]);
            if (lane == 0 && o) atomicOr(&s_or[j], o);
        }
    }
    const uint32_t my = (uint32_t)__cvta_generic_to_shared(sh[warp & (kSubHist - 1)]);
    if (my & 1023u) __trap();  // hist_inc needs 1 KB-aligned sub-histograms in the shared window (cannot happen: fail loudly)
    uint32_t zero_rows = 0;  // RGBA rows whose alpha residuals are all zero (an opaque plane): counted, not looked up
#pragma unroll
    for (int q = 0; q < kBH / kEncWarps; ++q) {
        if (C == 4 && __all_sync(0xFFFFFFFFu, nv[q] == 4 * C &&
                                 ((res[q][0] | res[q][1 % C] | res[q][2 % C] | res[q][3 % C]) & 0xFF000000u) == 0)) {
#pragma unroll
            for (int j = 0; j < 4 * C; ++j)
                if ((j & 3) != 3) hist_inc(my, res[q][j >> 2], j & 3);
            ++zero_rows;
        } else if (nv[q] == 4 * C) {
#pragma unroll
            for (int j = 0; j < 4 * C; ++j) hist_inc(my, res[q][j >> 2], j & 3);
        } else if (nv[q] > 0) {
#pragma unroll
            for (int j = 0; j < 4 * C; ++j)
                if (j < nv[q]) hist_inc(my, res[q][j >> 2], j & 3);
        }
    }
    if (C == 4 && lane == 0 && zero_rows) atomicAdd(&sh[warp & (kSubHist - 1)][0], zero_rows * (uint32_t)kBW);
    __syncthreads();
    if (kTma && resid && tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(resid + gb * (uint64_t)(kBH * 32 * C)),
                     "r"((uint32_t)__cvta_generic_to_shared(ptile)), "r"((uint32_t)(kBH * 128 * C)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kSubHist; ++k) s += sh[k][tid];
    {   // flat channels (FLP0 §2b): T = per channel byte, the OR of all its residual bytes bar the first pixel's
        uint32_t T;
        if (C == 4) T = s_or[0];
        else if (C == 3) {  // word j byte b carries channel (4j + b) mod 3
            const uint32_t a = s_or[0], b = s_or[1 % C], c = s_or[2 % C];
            T = (a | __byte_perm(b, 0u, 0x4102) | __byte_perm(c, 0u, 0x4021) | (a >> 24) |
                 __byte_perm(b, 0u, 0x4434) | __byte_perm(c, 0u, 0x4344)) & 0x00FFFFFFu;
        } else if (C == 2) { const uint32_t x = s_or[0] | s_or[1 % C]; T = (x | (x >> 16)) & 0xFFFFu; }
        else { uint32_t x = s_or[0]; x |= x >> 16; T = (x | (x >> 8)) & 0xFFu; }
        auto nonzero_bytes = [](uint32_t x) { x |= x >> 4; x |= x >> 2; x |= x >> 1; return x & 0x01010101u; };
        constexpr uint32_t chb = C == 4 ? 0x01010101u : ((1u << (8 * (C & 3))) - 1u) & 0x01010101u;
        const uint32_t flatb = ~nonzero_bytes(T) & chb;  // bit 8*ch: channel ch is flat
        const uint32_t first = s_first, npix = p.bwa * p.bha;
        if (tid == 0) s -= __popc(flatb) * (npix - 1u);   // all those residuals were zeros ...
        s -= __popc(~nonzero_bytes(first ^ ((uint32_t)tid * 0x01010101u)) & flatb);  // ... bar the first pixel's value
        if (flat && tid == 0) flat[gb] = make_uint2((flatb * 0x01020408u) >> 24, first & (flatb * 0xFFu));
    }
    hist[gb * 256 + tid] = (uint16_t)s;
    if (kTma && resid && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile must outlive the store's read of it
}

void launch_histograms(const uint8_t *d_pixels, const Geo &g, uint16_t *d_hist, uint32_t *d_resid, uint2 *d_flat,
                       const void *tensor_map, cudaStream_t s) {
    uint64_t total = (uint64_t)g.n * g.nb;
    const bool sg = (g.flags & FLIC_FLAG_SUBGREEN) && g.c >= 3;
    const bool g3 = grid3_ok(g);
    const dim3 grid = g3 ? dim3(g.nbx, g.nby, g.n) : dim3((unsigned)total);
    CUtensorMap tm;
    if (tensor_map) memcpy(&tm, tensor_map, sizeof tm); else memset(&tm, 0, sizeof tm);
#define FLIC_HIST(C, SG)                                                                                            \
    do {                                                                                                            \
        if (tensor_map) k_histograms<C, SG, true><<<grid, kEncThreads, 0, s>>>(d_pixels, g, d_hist, d_resid, d_flat, tm, g3);  \
        else k_histograms<C, SG, false><<<grid, kEncThreads, 0, s>>>(d_pixels, g, d_hist, d_resid, d_flat, tm, g3);            \
    } while (0)
    switch (g.c) {
        case 1: FLIC_HIST(1, false); break;
        case 2: FLIC_HIST(2, false); break;
        case 3: if (sg) FLIC_HIST(3, true); else FLIC_HIST(3, false); break;
        default: if (sg) FLIC_HIST(4, true); else FLIC_HIST(4, false); break;
    }
#undef FLIC_HIST
}

// -------------------------------------------------------------------- k_tables
// One warp owns up to 32 blocks.  Sorting and code assignment are warp-cooperative, one block at
// a time; the inherently serial Huffman merge runs with ONE BLOCK PER LANE (all 32 lanes busy) as
// the in-place three-pass minimum-redundancy construction of Moffat & Katajainen, whose tie rule
// (an internal node is taken only if strictly lighter than the next leaf) is FLP0 §3.2's.
constexpr int kTabPitch = 258;  // u16 elements per row: 129 words
constexpr int kTabBpw = 8;      // blocks per warp: 7.7 KB of smem per warp -> ~29 warps per SM
template <int BPW>
struct TabSmemT {
    uint32_t key[256];      // (count << 8) | symbol of the block being sorted
    uint16_t A[BPW * kTabPitch];  // merge arrays, one row per lane-owned block; odd word pitch = conflict-free in step
    uint8_t ord[BPW][256];   // sorted symbol order per block
    uint8_t lenS[256];      // code length per SYMBOL of the block being finished
    uint32_t next[16];      // first canonical code per length
};

// 12 nine-bit counters (lengths 1..12) packed into two u64
__device__ __forceinline__ void cnt_add(uint64_t &a, uint64_t &b, uint32_t l, uint32_t v = 1) {
    if (l >= 1 && l <= 6) a += (uint64_t)v << (9 * (l - 1));
    else if (l >= 7 && l <= 12) b += (uint64_t)v << (9 * (l - 7));
}
__device__ __forceinline__ void cnt_sub(uint64_t &a, uint64_t &b, uint32_t l) {
    if (l >= 1 && l <= 6) a -= 1ull << (9 * (l - 1));
    else if (l >= 7 && l <= 12) b -= 1ull << (9 * (l - 7));
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

// FLP0 §3.2-3.4 for one block held by one lane: sorted weights in A[i] (i < n, n >= 2)
// -> leaves per code length (packed counters), length-limited.
__device__ __forceinline__ void lane_lengths(uint16_t *A, int n, uint64_t &na, uint64_t &nb) {
#define AT(i) A[(i)]
    // pass 1: in-place merge; A[0..root) become parent pointers, A[root..next) internal weights
    AT(0) = (uint16_t)(AT(0) + AT(1));
    int root = 0, leaf = 2;
    for (int next = 1; next < n - 1; ++next) {
        uint32_t wsum;
        if (leaf >= n || AT(root) < AT(leaf)) { wsum = AT(root); AT(root++) = (uint16_t)next; }
        else wsum = AT(leaf++);
        if (leaf >= n || (root < next && AT(root) < AT(leaf))) { wsum += AT(root); AT(root++) = (uint16_t)next; }
        else wsum += AT(leaf++);
        AT(next) = (uint16_t)wsum;
    }
    // pass 2: parent pointers -> internal depths
    AT(n - 2) = 0;
    for (int next = n - 3; next >= 0; --next) AT(next) = (uint16_t)(AT(AT(next)) + 1);
    // pass 3: leaves per depth, depths beyond L folded into L
    na = nb = 0;
    int avbl = 1, dpth = 0;
    root = n - 2;
    while (avbl > 0) {
        int used = 0;
        while (root >= 0 && (int)AT(root) == dpth) { ++used; --root; }
        if (avbl > used) cnt_add(na, nb, (uint32_t)min(dpth, kL), (uint32_t)(avbl - used));
        avbl = 2 * used;
        ++dpth;
    }
#undef AT
    // Kraft repair
    uint32_t total = 0;
#pragma unroll
    for (int l = 1; l <= kL; ++l) total += cnt_get(na, nb, l) << (kL - l);
    while (total > (1u << kL)) {
        cnt_sub(na, nb, kL);
        for (int l = kL - 1; l >= 1; --l)
            if (cnt_get(na, nb, l)) { cnt_sub(na, nb, l); cnt_add(na, nb, l + 1, 2); break; }
        --total;
    }
}

// Bitonic sort of 32*K keys held K per lane (element e = lane*K + k), ascending: strides below K
// are register-to-register, the rest one shuffle per key.
template <int K>
__device__ __forceinline__ void warp_sort(uint32_t (&v)[K], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * K; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= K) {
                const int lj = j / K;
                const bool lower = (lane & lj) == 0;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, v[i], lj);
                    const bool asc = (((lane * K + i) & k) == 0);
                    v[i] = (lower == asc) ? min(v[i], o) : max(v[i], o);
                }
            } else {
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if ((i & j) == 0) {
                        const bool asc = (((lane * K + i) & k) == 0);
                        const uint32_t a = v[i], b = v[i | j];
                        v[i] = asc ? min(a, b) : max(a, b);
                        v[i | j] = asc ? max(a, b) : min(a, b);
                    }
                }
            }
        }
    }
}

// Sorts key[0..32K) in shared memory through registers and scatters (weight, symbol) of the first n.
template <int K>
__device__ __forceinline__ void sort_block(uint32_t *key, int n, uint16_t *Arow, uint8_t *ord, int lane) {
    uint32_t v[K];
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = key[lane * K + i];
    warp_sort<K>(v, lane);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int e = lane * K + i;
        if (e < n) { Arow[e] = (uint16_t)(v[i] >> 8); ord[e] = (uint8_t)v[i]; }
    }
}

// bits (optional): per block, the sum over symbols of count x code length — with the geometry that is
// the block's slot size (FLP0 §7), so every block's output position is known before k_pack runs.
// kTabWarps warps per CTA, each on its own run of blocks (they share nothing, there is no CTA barrier): one-warp CTAs ran
// into the SM's limit of 32 resident CTAs and their launch overhead (17 warps resident on average of the 29 that fit).
#ifndef FLIC_TAB_WARPS
#define FLIC_TAB_WARPS 4
#endif
constexpr int kTabWarps = FLIC_TAB_WARPS;
template <int BPW>
__global__ void __launch_bounds__(32 * kTabWarps) k_tables(const uint16_t *__restrict__ hist, uint64_t nblocks,
                                               uint16_t *__restrict__ table, uint32_t *__restrict__ bits, int bpw) {
    __shared__ __align__(16) TabSmemT<BPW> sw[kTabWarps];
    const int lane = threadIdx.x & 31;
    TabSmemT<BPW> &s = sw[threadIdx.x >> 5];
    const uint64_t first = ((uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * bpw;
    if (first >= nblocks) return;
    const int cnt = (int)min((uint64_t)bpw, nblocks - first);
    int myn = 0;

    // ---- cooperative: compact + sort the used symbols of each block (FLP0 §3.1)
    for (int j = 0; j < cnt; ++j) {
        const uint4 hv = *reinterpret_cast<const uint4 *>(hist + (first + j) * 256 + 8 * lane);
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
        uint32_t f[8], nact = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { f[k] = (hw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu; nact += f[k] != 0; }
        const uint32_t incl = warp_incl_scan(nact, lane);
        const int n = (int)__shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t pos = incl - nact;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (f[k]) s.key[pos++] = (f[k] << 8) | (uint32_t)(8 * lane + k);
        const int P = n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256));
        for (int i = n + lane; i < P; i += 32) s.key[i] = 0xFFFFFFFFu;
        __syncwarp();
        uint16_t *Arow = s.A + j * kTabPitch;
        if (P == 32) sort_block<1>(s.key, n, Arow, s.ord[j], lane);
        else if (P == 64) sort_block<2>(s.key, n, Arow, s.ord[j], lane);
        else if (P == 128) sort_block<4>(s.key, n, Arow, s.ord[j], lane);
        else sort_block<8>(s.key, n, Arow, s.ord[j], lane);
        if (lane == j) myn = n;
        __syncwarp();
    }

    // ---- one block per lane: merge, depths, census, Kraft repair (FLP0 §3.2-3.4)
    uint64_t mya = 0, myb = 0;
    if (lane < cnt && myn >= 2) lane_lengths(s.A + lane * kTabPitch, myn, mya, myb);
    __syncwarp();

    // ---- cooperative: lengths by rank, canonical codes, table out (FLP0 §3.5, §4)
    for (int j = 0; j < cnt; ++j) {
        const int n = __shfl_sync(0xFFFFFFFFu, myn, j);
        const uint64_t na = shfl64(mya, j), nb = shfl64(myb, j);
        reinterpret_cast<uint2 *>(s.lenS)[lane] = make_uint2(0u, 0u);
        uint32_t nx = 0, prevnum = 0;
        uint32_t thr[kL + 2];  // thr[l] = number of ranks whose code is at least l bits long (the rarest ranks get the longest codes)
#pragma unroll
        for (int l = 1; l <= kL; ++l) {
            nx = (nx + prevnum) << 1;
            if (lane == 0) s.next[l] = nx;
            prevnum = cnt_get(na, nb, l);
            thr[l] = prevnum;
        }
        thr[kL + 1] = 0;
#pragma unroll
        for (int l = kL; l >= 1; --l) thr[l] += thr[l + 1];
        __syncwarp();
        if (n >= 2) {
            // rank i (ascending weight) has length kL - #{t in 2..kL : i >= thr[t]}: ten compares instead of ten fill loops
            for (int i = lane; i < n; i += 32) {
                uint32_t l = (uint32_t)kL;
#pragma unroll
                for (int t = kL; t >= 2; --t) l -= (uint32_t)i >= thr[t] ? 1u : 0u;
                s.lenS[s.ord[j][i]] = (uint8_t)l;
            }
        } else if (n == 1 && lane == 0) {
            s.lenS[s.ord[j][0]] = (uint8_t)kLenSole;
        }
        __syncwarp();
        // canonical codes in (length, symbol) order; the lane owns symbols 8*lane..8*lane+7
        const uint2 lw = reinterpret_cast<const uint2 *>(s.lenS)[lane];
        uint32_t l8[8];
        uint64_t ca = 0, cb = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            l8[k] = ((k < 4 ? lw.x : lw.y) >> (8 * (k & 3))) & 0xFFu;
            cnt_add(ca, cb, l8[k]);
        }
        uint64_t ia = ca, ib = cb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t ta = shfl_up64(ia, d), tb = shfl_up64(ib, d);
            if (lane >= d) { ia += ta; ib += tb; }
        }
        uint64_t ea = ia - ca, eb = ib - cb;  // symbols of each length in lower lanes
        if (bits) {
            const uint4 hv = *reinterpret_cast<const uint4 *>(hist + (first + j) * 256 + 8 * lane);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
            uint32_t b = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (l8[k] <= (uint32_t)kL) b += ((hw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu) * l8[k];
            b = __reduce_add_sync(0xFFFFFFFFu, b);
            if (lane == 0) bits[first + j] = b;
        }
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
        *reinterpret_cast<uint4 *>(table + (first + j) * 256 + 8 * lane) = make_uint4(out[0], out[1], out[2], out[3]);
        __syncwarp();
    }
}

void launch_tables(const uint16_t *d_hist, uint64_t nblocks, uint16_t *d_table, uint32_t *d_bits, cudaStream_t s) {
    // one block per warp for small jobs (latency: 512x512 RGB 27 -> 21 us), a few when that still fills the
    // chip, up to kTabBpw blocks per warp (merge lanes busy) for large jobs (measured: 2 / 4 / 16 per warp are slower)
    uint64_t want = nblocks / (148ull * 16);
    int bpw = (int)(want < 1 ? 1 : (want > (uint64_t)kTabBpw ? kTabBpw : want));
    const uint64_t warps = (nblocks + bpw - 1) / bpw;
    const int wpc = warps >= 148ull * 4 * kTabWarps ? kTabWarps : 1;  // small jobs: one warp per CTA spreads over more SMs
    unsigned grid = (unsigned)((warps + wpc - 1) / wpc);
    k_tables<kTabBpw><<<grid, 32 * wpc, 0, s>>>(d_hist, nblocks, d_table, d_bits, bpw);
}

// ------------------------------------------------------------------ k_finalize
// Headers, rebased u32 directories and the n+1 stream offsets: all of it follows from dirE, so it needs
// only k_slots' result (not k_pack's).  `first`/`stride` spread the items over whoever calls this.
__device__ __forceinline__ void finalize_items(const Geo &g, const unsigned long long *dirE, uint32_t *streams,
                                               uint64_t capacity_words, unsigned long long *offsets, uint32_t *err,
                                               uint64_t first_item, uint64_t stride) {
    const uint64_t per = kHdrWords + (uint64_t)g.nb + 1;  // header + directory words per image
    const uint64_t items = (uint64_t)g.n * per;
    for (uint64_t i = first_item; i < items + g.n + 1; i += stride) {
        if (i >= items) {  // stream offsets in bytes
            uint64_t img = i - items;
            offsets[img] = 4ull * (img * per + dirE[img * g.nb]);
            continue;
        }
        uint64_t img = i / per, k = i - img * per;
        unsigned long long first = dirE[img * g.nb];
        uint64_t pos = img * per + first + k;
        if (pos >= capacity_words) continue;  // k_pack raises kErrCapacity
        uint32_t v;
        if (k >= kHdrWords) {
            v = (uint32_t)(dirE[img * g.nb + (k - kHdrWords)] - first);
        } else {
            switch (k) {
                case 0: v = kMagic; break;
                case 1: v = kVersion | (g.c << 16) | ((g.flags & 0xFFu) << 24); break;
                case 2: v = g.w; break;
                case 3: v = g.h; break;
                case 4: v = (uint32_t)kBW | ((uint32_t)kBH << 16); break;
                case 5: v = g.nb; break;
                case 6: {  // payload words: the container's offsets are u32
                    const unsigned long long pw = dirE[(img + 1) * g.nb] - first;
                    if (pw > 0xFFFFFFFFull) atomicOr(err, kErrRange);
                    v = (uint32_t)pw;
                    break;
                }
                default: v = (uint32_t)kL; break;
            }
        }
        streams[pos] = v;
    }
}

__global__ void __launch_bounds__(256) k_finalize(Geo g, const unsigned long long *__restrict__ dirE,
                                                  uint32_t *__restrict__ streams, uint64_t capacity_words,
                                                  unsigned long long *__restrict__ offsets, uint32_t *err) {
    finalize_items(g, dirE, streams, capacity_words, offsets, err, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x,
                   (uint64_t)gridDim.x * blockDim.x);
}

// ---------------------------------------------------------------------- k_slots
// FLP0 §7: slot(block) = block header + floor(code bits / 32) + one word per real row (none without code
// bits); dirE = exclusive prefix sum over all blocks of the batch (dirE[total] = grand total).
// Half a megabyte of input, so latency is all that matters: at most 128 co-resident CTAs each own a
// contiguous run of blocks, publish the run's sum tagged with this launch's epoch (no memset between
// launches), sum their predecessors' published values, then scan their own run.
constexpr int kSlotThreads = 1024;
// one: FLIC_FLAG_ONE_STREAM — the block is its header and ONE bit stream, whose exact length the histogram gives.
__device__ __forceinline__ uint32_t slot_words(uint32_t code_bits, uint32_t b, uint32_t last_row_first, uint32_t last_bha, bool one) {
    if (one) return (uint32_t)kBlkHdrWords1 + ((code_bits + 31u) >> 5);
    return (uint32_t)kBlkHdrWords + (code_bits >> 5) + (code_bits ? (b >= last_row_first ? last_bha : (uint32_t)kBH) : 0u);
}
__global__ void __launch_bounds__(kSlotThreads) k_slots(Geo g, const uint32_t *__restrict__ bits,
                                                        unsigned long long *__restrict__ dirE, uint32_t per,
                                                        uint32_t epoch, unsigned long long *status,
                                                        uint64_t capacity_words, uint32_t *err, uint32_t *streams,
                                                        unsigned long long *offsets) {
    __shared__ uint32_t wsum[32];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t total = (uint64_t)g.n * g.nb;
    const uint64_t lo = (uint64_t)blockIdx.x * per, hi = min(total, lo + per);
    const uint32_t last_row_first = (g.nby - 1) * g.nbx, last_bha = g.h - (g.nby - 1) * kBH;
    const bool one = one_stream(g.flags);
    // pass 1: the run's sum
    uint32_t sum = 0;
    for (uint64_t gb = lo + tid; gb < hi; gb += kSlotThreads)
        sum += slot_words(__ldg(bits + gb), (uint32_t)(gb % g.nb), last_row_first, last_bha, one);
    sum = __reduce_add_sync(0xFFFFFFFFu, sum);
    if (lane == 0) wsum[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        const uint32_t agg = __reduce_add_sync(0xFFFFFFFFu, wsum[lane]);
        if (lane == 0) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(status + blockIdx.x),
                         "l"(((unsigned long long)epoch << 32) | agg) : "memory");
        }
        unsigned long long excl = 0;
        for (uint32_t i = lane; i < blockIdx.x; i += 32) {
            unsigned long long v;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(status + i) : "memory");
            } while ((uint32_t)(v >> 32) != epoch);
            excl += (uint32_t)v;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const uint32_t l = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)excl, d), h = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(excl >> 32), d);
            excl += ((unsigned long long)h << 32) | l;
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    // pass 2: exclusive scan of the run, 1024 blocks per step (the second read of `bits` hits L2)
    unsigned long long carry = s_excl;
    for (uint64_t base = lo; base < hi; base += kSlotThreads) {
        const uint64_t gb = base + tid;
        const uint32_t v = gb < hi ? slot_words(__ldg(bits + gb), (uint32_t)(gb % g.nb), last_row_first, last_bha, one) : 0u;
        const uint32_t incl = warp_incl_scan(v, lane);
        __syncthreads();
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        const uint32_t w = wsum[lane];
        const uint32_t wincl = warp_incl_scan(w, lane);
        const uint32_t before = __shfl_sync(0xFFFFFFFFu, wincl - w, warp);
        if (gb < hi) dirE[gb] = carry + before + (incl - v);
        carry += __shfl_sync(0xFFFFFFFFu, wincl, 31);
    }
    if (hi == total && tid == 0) {
        dirE[total] = carry;
        if ((uint64_t)g.n * (kHdrWords + (uint64_t)g.nb + 1) + carry > capacity_words) atomicOr(err, kErrCapacity);
    }
    if (gridDim.x == 1 && streams) {  // small job, one CTA: write the headers and directories here, save a launch
        __syncthreads();
        finalize_items(g, dirE, streams, capacity_words, offsets, err, tid, kSlotThreads);
    }
}

// CTAs of k_slots spin on their predecessors' published sums, so every CTA of the grid must be resident at once:
// the bound comes from the occupancy API, not from an assumption about the device.
int slots_max_resident_ctas() {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_slots, kSlotThreads, 0) != cudaSuccess)
        return 0;
    const int n = sms * per_sm;
    return n > 128 ? 128 : n;  // the status array holds 128 run sums
}

// status: >= 128 u64, zeroed once at allocation; epoch: a value never used before on this status array (>= 1)
// max_grid: slots_max_resident_ctas() of the device (1..128).
// Returns true when the (single) CTA also wrote headers and directories, i.e. k_finalize is not needed.
bool launch_slots(const Geo &g, const uint32_t *d_bits, unsigned long long *d_dirE, unsigned long long *d_status,
                  uint32_t epoch, uint64_t capacity_words, uint32_t *d_err, uint32_t *d_streams,
                  unsigned long long *d_offsets, int max_grid, cudaStream_t s) {
    const uint64_t total = (uint64_t)g.n * g.nb;
    uint64_t per = (total + max_grid - 1) / max_grid;
    per = ((per < 2048 ? 2048 : per) + kSlotThreads - 1) / kSlotThreads * kSlotThreads;
    const unsigned grid = (unsigned)((total + per - 1) / per);  // <= max_grid: all CTAs are resident, the spin cannot deadlock
    k_slots<<<grid, kSlotThreads, 0, s>>>(g, d_bits, d_dirE, (uint32_t)per, epoch, d_status, capacity_words, d_err, d_streams,
                                          d_offsets);
    return grid == 1;
}

// ------------------------------------------------------- decoupled look-back (k_pack EXACT, k_encode)
// look-back status word: epoch (22 bits) | flag (2) | value (40 bits, words)
constexpr unsigned long long kStA = 1ull << 40, kStP = 2ull << 40, kStVal = (1ull << 40) - 1;
// A status word IS the message (epoch, flag and value travel in one 64-bit store; nothing else in memory is published
// by it), so relaxed accesses at GPU scope are all the ordering it needs.  With st.release / ld.acquire every polling
// load carried a fence: k_pack EXACT 1.67 ms instead of 1.48 on the 4K RGBA batch.
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// One warp: the exclusive prefix of the sizes of all blocks before gb (gb > 0), 32 predecessors per step.  A predecessor
// publishes its own size (kStA) as soon as it knows it and its inclusive prefix (kStP) once it has looked back itself.
// Block indices are handed out in order (a ticket counter), so every predecessor is running or done: the wait ends.
__device__ __forceinline__ unsigned long long lookback_excl(const unsigned long long *status, uint64_t gb, uint32_t epoch, int lane) {
    // One window of 32 per step.  Everything wider was measured SLOWER on the 4K RGBA batch (k_pack EXACT, ms): four
    // windows requested at once 1.67 -> 1.84; the whole CTA looking back (256 per step) 1.48 -> 1.72; a size pass that
    // publishes the block's size before it is packed (so that nobody waits for a straggler) 1.48 -> 1.92.  What the
    // look-back costs is status traffic and the instructions around it, not the depth of the search: with the positions
    // of an identical earlier launch in place of the look-back the kernel takes 1.10 ms.
    constexpr int kLbK = 1;
    unsigned long long excl = 0;
    long long j0 = (long long)gb - 1 - lane;
    for (;;) {
        unsigned long long v[kLbK];
#pragma unroll
        for (int k = 0; k < kLbK; ++k) v[k] = j0 - 32 * k >= 0 ? ld_status(status + (j0 - 32 * k)) : 0ull;
        unsigned long long c = 0;
        bool done = false;
#pragma unroll
        for (int k = 0; k < kLbK; ++k) {
            const long long j = j0 - 32 * k;
            if (!done) {
                if (j >= 0) {
                    // a predecessor that has not published yet is still working on its block: sleep instead of spinning,
                    // the issue slots are what the other CTAs of the SM are short of
                    while ((v[k] >> 42) != epoch || (v[k] & (kStA | kStP)) == 0) { __nanosleep(64); v[k] = ld_status(status + j); }
                }
                const uint32_t pmask = __ballot_sync(0xFFFFFFFFu, j >= 0 && (v[k] & kStP) != 0);
                const int stop = pmask ? __ffs(pmask) - 1 : 32;  // nearest predecessor that already knows its inclusive prefix
                if (j >= 0 && lane <= stop) c += v[k] & kStVal;
                if (pmask || j - (31 - lane) <= 0) done = true;  // found a prefix, or the window reached block 0 (uniform: lane 31's j)
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const uint32_t lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)c, d), hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(c >> 32), d);
            c += ((unsigned long long)hi << 32) | lo;
        }
        excl += c;
        if (done) break;
        j0 -= 32 * kLbK;
    }
    return excl;
}

// ------------------------------------------------- ONE_STREAM copy-out (k_pack kLayOne, k_encode)
// FLP0 §8: the rows back to back, bit-exactly, into out[0, nw).  Row r is staged MSB-first from word 0 of its staging
// row (zero past its last bit, one word beyond included) and starts at stream bit
// rbit0[r].  One warp per row: every lane shifts by the same amount, interior words of a row go straight out
// (coalesced); the first and the last word of a row may be shared with its neighbours, so they are parked as
// (word index, value) pairs — two per row, in stream order — and warp 0 ORs the runs of equal index together at the end.
struct EdgeWords { uint32_t idx[2 * kBH], val[2 * kBH]; };
__device__ __forceinline__ void concat_rows(const uint32_t *stage0 /* row 0, word 0 */, int pitch, const uint32_t *rbit0, EdgeWords &e,
                                            uint32_t *out, uint32_t nw, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < kBH; r += kEncWarps) {
        const uint32_t s0 = rbit0[r], s1 = rbit0[r + 1], sh = s0 & 31u, f = s0 >> 5;
        const uint32_t *src = stage0 + r * pitch;
        if (s1 == s0) {  // no bits: a zero contribution that keeps the neighbours' shared word adjacent in the list
            if (lane < 2) { e.idx[2 * r + lane] = f; e.val[2 * r + lane] = 0u; }
            continue;
        }
        const uint32_t l = (s1 - 1u) >> 5, nrw = l - f + 1u;
        for (uint32_t k = lane; k < nrw; k += 32) {
            const uint32_t v = __funnelshift_r(src[k], k ? src[k - 1] : 0u, sh);  // (src[k-1] << (32 - sh)) | (src[k] >> sh)
            if (k == 0) { e.idx[2 * r] = f; e.val[2 * r] = v; }
            else if (k == nrw - 1u) { e.idx[2 * r + 1] = l; e.val[2 * r + 1] = v; }
            else out[f + k] = v;
        }
        if (nrw == 1u && lane == 0) { e.idx[2 * r + 1] = l; e.val[2 * r + 1] = 0u; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            const uint32_t w = e.idx[i];
            if ((i == 0 || e.idx[i - 1] != w) && w < nw) {  // head of a run of equal word indices
                uint32_t v = e.val[i];
                for (int t = i + 1; t < 2 * kBH && e.idx[t] == w; ++t) v |= e.val[t];
                out[w] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------- k_pack
// Staging tile: per row 3 pad words (they absorb the all-zero upper words of a code group placed at
// the very start of a row) + kRowWordsMax data words; the pitch is odd, so the interleaving copy-out
// reads a column conflict-free.
constexpr int kStagePad = 3;
constexpr int kStagePitch = kStagePad + kRowWordsMax + ((kStagePad + kRowWordsMax) % 2 == 0 ? 1 : 0);
static_assert(kBH * kStagePitch % 4 == 0, "zero-fill uses 16-byte stores");

__device__ __forceinline__ void red_or_shared(uint32_t addr, int off, uint32_t v) {
    if (off == -4) asm volatile("red.shared.or.b32 [%0+-4], %1;" ::"r"(addr), "r"(v) : "memory");
    else if (off == -8) asm volatile("red.shared.or.b32 [%0+-8], %1;" ::"r"(addr), "r"(v) : "memory");
    else asm volatile("red.shared.or.b32 [%0+-12], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Opaque multipliers (kernel arguments): x >> k is issued as IMAD.HI(x, 2^(32-k)) on the FMA pipe instead
// of SHF on the ALU pipe, which is the one this kernel saturates.
struct PackMul { uint32_t m8, m10, m18, m26; };  // 2^8 (>>24), 2^10 (>>22), 2^18 (>>14), 2^26 (>>6)

// The code group ("quad") of the four symbols in word w, MSB-first: value in qhi:qlo (at most 40 bits,
// right-aligned), bit count returned.  A table entry is code | len << 24; the pair merge
// ((e0 << len1) | e1) & 0xFFFFF leaves the length fields above bit 24 where they are masked off, and the
// sum of entries carries the sum of lengths in its top byte (the code fields cannot carry into it).
// Bit b of the skip mask says byte b belongs to a flat channel (FLP0 §2b) and contributes nothing: SK >= 0
// is that mask at compile time (0, or 8 = the alpha byte of an RGBA word: no table read at all for it),
// SK < 0 takes it from `skip` at run time.
template <bool kFull, int SK>
__device__ __forceinline__ uint32_t quad_of(uint32_t w, int first, int nv, uint32_t skip, const uint32_t *tab,
                                            const PackMul &pm, uint32_t &qlo, uint32_t &qhi) {
    const char *t = reinterpret_cast<const char *>(tab);
    if (SK >= 0) skip = (uint32_t)SK;
    uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    if (!(skip & 1u)) e0 = *reinterpret_cast<const uint32_t *>(t + ((w << 2) & 0x3FCu));
    if (!(skip & 2u)) e1 = *reinterpret_cast<const uint32_t *>(t + (__umulhi(w, pm.m26) & 0x3FCu));
    if (!(skip & 4u)) e2 = *reinterpret_cast<const uint32_t *>(t + (__umulhi(w, pm.m18) & 0x3FCu));
    if (!(skip & 8u)) e3 = *reinterpret_cast<const uint32_t *>(t + (__umulhi(w, pm.m10) & 0x3FCu));
    if (!kFull) {  // ragged right edge: symbols at or past nv contribute nothing
        if (first + 0 >= nv) e0 = 0;
        if (first + 1 >= nv) e1 = 0;
        if (first + 2 >= nv) e2 = 0;
        if (first + 3 >= nv) e3 = 0;
    }
    const uint32_t pa = ((e0 << __umulhi(e1, pm.m8)) | e1) & 0xFFFFFu;
    const uint32_t pb = ((e2 << __umulhi(e3, pm.m8)) | e3) & 0xFFFFFu;
    const uint32_t sb = e2 + e3;
    const uint32_t lb = __umulhi(sb, pm.m8);
    qlo = (pa << lb) | pb;
    qhi = __funnelshift_l(pa, 0u, lb);
    return __umulhi(e0 + e1 + sb, pm.m8);
}

// The opaque-alpha RGBA word: three symbols, at most 3 * kL = 30 bits — the group is ONE register, it can only
// touch the word it ends in and the one before (two RED.OR, no "longer than 32 bits?" vote).
static_assert(3 * kL <= 32, "quad3 returns the group in one word");
__device__ __forceinline__ uint32_t quad3(uint32_t w, const uint32_t *tab, const PackMul &pm, uint32_t &q) {
    const char *t = reinterpret_cast<const char *>(tab);
    const uint32_t e0 = *reinterpret_cast<const uint32_t *>(t + ((w << 2) & 0x3FCu));
    const uint32_t e1 = *reinterpret_cast<const uint32_t *>(t + (__umulhi(w, pm.m26) & 0x3FCu));
    const uint32_t e2 = *reinterpret_cast<const uint32_t *>(t + (__umulhi(w, pm.m18) & 0x3FCu));
    const uint32_t pa = ((e0 << __umulhi(e1, pm.m8)) | e1) & 0xFFFFFu;
    q = (pa << __umulhi(e2, pm.m8)) | (e2 & 0xFFFFFFu);
    return __umulhi(e0 + e1 + e2, pm.m8);
}

// Layouts (FLP0 §3.8): where a block goes and what its payload looks like.
//   kLaySlots : position and size from k_slots (histogram x code lengths); interleaved row sub-streams; slack zeroed
//   kLayOne   : FLIC_FLAG_ONE_STREAM — position from k_slots too (the block's size is exactly header + ceil(code bits / 32));
//               the rows leave concatenated bit-exactly into ONE stream
//   kLayExact : FLIC_FLAG_EXACT — the block occupies exactly the words it packed, so its position is known only after the
//               packing: blocks are claimed in order (a ticket), publish their size and look back over their predecessors'
enum { kLaySlots = 0, kLayOne = 1, kLayExact = 2 };

template <int C, int LAY>
__global__ void __launch_bounds__(kEncThreads, 8) k_pack(const uint32_t *__restrict__ resid, Geo g,
                                                      const uint16_t *__restrict__ table,
                                                      const uint2 *__restrict__ flat,
                                                      uint32_t *__restrict__ streams, uint64_t capacity_words,
                                                      unsigned long long *dirE,
                                                      uint32_t *err, PackMul pm, unsigned long long *status,
                                                      unsigned long long *ticket, unsigned long long ticket_base, uint32_t epoch,
                                                      int grid3, const unsigned long long *__restrict__ part_base, uint32_t part_hdr_words) {
    __shared__ __align__(16) uint32_t stage[kBH * kStagePitch + 4];  // + 4: kLayOne reads one word past a row's last
    __shared__ uint32_t tab[256];  // code | len << 24; 0 for a sole symbol (no bits)
    __shared__ uint8_t nib[256];
    __shared__ uint32_t rwc[kBH], rowoff[kBH], rbit0[LAY == kLayOne ? kBH + 1 : 1];
    __shared__ EdgeWords edges;  // (kLayOne only; 512 bytes)
    __shared__ uint32_t s_minw, s_used;
    __shared__ unsigned long long s_excl, s_ticket;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t gb;
    BlockPos p;
    // blocks in ticket order (EXACT): whoever holds block gb knows that every block before it has started.  The ticket is
    // drawn first and read after the zero-fill, which does not depend on it and covers the atomic's latency.
    if (LAY == kLayExact && tid == 0) s_ticket = atomicAdd(ticket, 1ull) - ticket_base;
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *s4 = reinterpret_cast<uint4 *>(stage);
        for (int i = tid; i < (kBH * kStagePitch + 4) / 4; i += kEncThreads) s4[i] = z;
    }
    if (LAY == kLayExact) {
        __syncthreads();
        gb = s_ticket;
        p = block_pos(g, gb);
    } else {
        p = block_pos_cta(g, grid3 != 0, gb);
    }

    {
        uint32_t e = table[gb * 256 + tid], l = e >> 12;
        nib[tid] = (uint8_t)l;
        tab[tid] = (l == kLenSole || l == 0) ? 0u : ((e & 0xFFFu) | (l << 24));
    }
    // residual words of the lane's four pixels in row r (k_histograms wrote them in this lane order)
    auto load_row = [&](int r, uint32_t(&v)[C]) {
        const uint32_t *t = resid + ((gb * kBH + r) * 32 + lane) * C;
        if (C == 4) {
            uint4 x = ldg_nc_v4(t);
            v[0] = x.x; v[1 % C] = x.y; v[2 % C] = x.z; v[3 % C] = x.w;
        } else if (C == 2) {
            uint2 x = __ldg(reinterpret_cast<const uint2 *>(t));
            v[0] = x.x; v[1 % C] = x.y;
        } else {
#pragma unroll
            for (int j = 0; j < C; ++j) v[j] = __ldg(t + j);
        }
    };
    const int nvfull = C * max(0, min(4, (int)p.bwa - 4 * lane));
    const uint2 fl = flat[gb];  // {mask of flat channels, their values}: CTA-uniform
    uint32_t skip[C];
#pragma unroll
    for (int j = 0; j < C; ++j) skip[j] = word_channel_bits<C>(fl.x, j);
    uint32_t nxt[C];
    load_row(warp, nxt);
    __syncthreads();

    // FLP0 §5: one warp per row.  Each lane turns its 4*C symbols into C code groups, a warp scan of bit
    // counts places them, and every group is OR-ed straight into the zeroed staging row at its END bit
    // position: group << ((32 - end) & 31) occupies the word the group ends in and the one or two before.
    const uint32_t sstage = (uint32_t)__cvta_generic_to_shared(stage);
    // kNarrow: an opaque-alpha RGBA block (flat mask 8) — every group is at most 30 bits (quad3, or the generic builder with
    // the alpha byte skipped on ragged lanes) and fits one register.  Two instances of the whole row loop, chosen once per
    // block (1.156 -> 1.065 ms on the 4K RGBA batch).  Measured side effect, unexplained by the SASS (the wide instance's
    // hot path is instruction-for-instruction the one of a build without the narrow instance): blocks that take the wide
    // instance pack 4 % slower than in such a build (noise batch 1.755 -> 1.83 ms); neither the order of the two
    // instances, a smaller unroll, nor moving the rare ragged rows out of line changed that.
    auto pack_rows = [&](auto narrow_c) {
        constexpr bool kNarrow = decltype(narrow_c)::value;
#pragma unroll
        for (int q = 0; q < kBH / kEncWarps; ++q) {
            const int r = warp + kEncWarps * q;
            uint32_t cur[C];
#pragma unroll
            for (int j = 0; j < C; ++j) cur[j] = nxt[j];
            if (q + 1 < kBH / kEncWarps) load_row(r + kEncWarps, nxt);  // next row's residuals fly while this one packs
            const int nv = r < (int)p.bha ? nvfull : 0;
            uint32_t qlo[C], qhi[C], ql[C], nbits = 0;
            if (kNarrow && nv == 4 * C) {  // three table reads per pixel
#pragma unroll
                for (int j = 0; j < C; ++j) { ql[j] = quad3(cur[j], tab, pm, qlo[j]); nbits += ql[j]; }
            } else if (!kNarrow && nv == 4 * C && fl.x == 0) {
#pragma unroll
                for (int j = 0; j < C; ++j) { ql[j] = quad_of<true, 0>(cur[j], 4 * j, 4 * C, 0u, tab, pm, qlo[j], qhi[j]); nbits += ql[j]; }
            } else {
#pragma unroll
                for (int j = 0; j < C; ++j) { ql[j] = quad_of<false, -1>(cur[j], 4 * j, nv, skip[j], tab, pm, qlo[j], qhi[j]); nbits += ql[j]; }
            }
            const uint32_t incl = warp_incl_scan(nbits, lane);
            if (lane == 31) {
                rwc[r] = (incl + 31u) >> 5;
                if (LAY == kLayOne) rbit0[r] = incl;
            }
            // nb = -(bit address in shared memory of the lane's next free bit); its low 5 bits are the shift
            uint32_t nb = 0u - (8u * (sstage + 4u * (uint32_t)(r * kStagePitch + kStagePad)) + (incl - nbits));
            if (kNarrow) {
                // What bounds this kernel is the LSU data pipe (ncu: 77 % of peak, a quarter of it these RED.ORs at ~2.4
                // wavefronts each), so two groups are first joined into one string of at most 60 bits: two RED.ORs per
                // PAIR, and a third only from the lanes whose string reaches a third word.
#pragma unroll
                for (int j = 0; j + 1 < C; j += 2) {
                    const uint32_t l1 = ql[j + 1];
                    const uint32_t plo = (qlo[j] << l1) | qlo[j + 1];        // l1 <= 30
                    const uint32_t phi = __funnelshift_l(qlo[j], 0u, l1);    // qlo[j] >> (32 - l1); 0 when l1 == 0
                    nb -= ql[j] + l1;
                    const uint32_t a = ((31u - nb) >> 3) & ~3u;  // byte address one past the word the string ends in
                    red_or_shared(a, -4, __funnelshift_l(0u, plo, nb));
                    red_or_shared(a, -8, __funnelshift_l(plo, phi, nb));
                    const uint32_t w2 = __funnelshift_l(phi, 0u, nb);
                    if (w2) red_or_shared(a, -12, w2);
                }
            } else {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    nb -= ql[j];
                    const uint32_t a = ((31u - nb) >> 3) & ~3u;  // byte address one past the word the group ends in
                    red_or_shared(a, -4, __funnelshift_l(0u, qlo[j], nb));
                    red_or_shared(a, -8, __funnelshift_l(qlo[j], qhi[j], nb));
                    if (__any_sync(0xFFFFFFFFu, ql[j] > 32u)) red_or_shared(a, -12, __funnelshift_l(qhi[j], 0u, nb));
                }
            }
        }
    };
    if (C == 4 && fl.x == 8u) pack_rows(std::true_type{});
    else pack_rows(std::false_type{});
    __syncthreads();

    constexpr int hdrw = LAY == kLayOne ? kBlkHdrWords1 : kBlkHdrWords;
    unsigned long long excl = 0;
    uint32_t slot = 0;
    if (LAY != kLayExact) {
        // FLP0 §7: the block's slot (position and size) was fixed by k_slots from the histogram and the code
        // lengths, so there is nothing to wait for: no ordering between blocks, no look-back.
        excl = dirE[gb];
        slot = (uint32_t)(dirE[gb + 1] - excl);
    }
    if (warp == 0) {
        uint32_t used, mn = 0;
        if (LAY == kLayOne) {  // first bit of every row in the block's one stream
            const uint32_t rb = rbit0[lane];
            const uint32_t bincl = warp_incl_scan(rb, lane);
            __syncwarp();
            rbit0[lane] = bincl - rb;
            if (lane == 31) rbit0[kBH] = bincl;
            used = (uint32_t)hdrw + ((__shfl_sync(0xFFFFFFFFu, bincl, 31) + 31u) >> 5);
        } else {
            const uint32_t wcount = rwc[lane];
            const uint32_t incl = warp_incl_scan(wcount, lane);
            rowoff[lane] = incl - wcount;
            mn = lane < (int)p.bha ? wcount : 0xFFFFFFFFu;  // FLP0 §6: interleave depth
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
            used = (uint32_t)hdrw + __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (LAY == kLayExact) {
            slot = used;
            if (gb > 0) {
                if (lane == 0) st_status(status + gb, ((unsigned long long)epoch << 42) | kStA | slot);
                excl = lookback_excl(status, gb, epoch, lane);
            }
            if (lane == 0) {
                st_status(status + gb, ((unsigned long long)epoch << 42) | kStP | ((excl + slot) & kStVal));
                dirE[gb] = excl;
                if (gb + 1 == (uint64_t)g.n * g.nb) dirE[gb + 1] = excl + slot;
                s_excl = excl;
            }
        }
        if (lane == 0) {
            s_minw = mn;
            s_used = used;
            // cannot happen: a slot covers the rows' padding (slots), is the stream's exact length (one stream, exact)
            if (used > slot || (LAY != kLaySlots && used != slot)) atomicOr(err, kErrSlot);
        }
    }
    __syncthreads();
    if (LAY == kLayExact) { excl = s_excl; slot = s_used; }
    if (s_used > slot || (LAY != kLaySlots && s_used != slot)) return;
    // part_base (block-row split across GPUs, api.cu flic_encode_emit_device): this launch's blocks are a run of block rows of
    // a larger image whose spliced stream starts at `streams` — possibly in another GPU's memory, written over NVLink —
    // with part_hdr_words of header + directory in front of the payload and *part_base payload words of earlier parts.
    const unsigned long long base = part_base ? (unsigned long long)part_hdr_words + *part_base + excl
                                              : (unsigned long long)(p.img + 1) * (kHdrWords + g.nb + 1) + excl;
    if (base + slot > capacity_words) {
        if (tid == 0) atomicOr(err, kErrCapacity);
        return;
    }
    uint32_t *out = streams + base;
    for (uint32_t i = s_used + tid; i < slot; i += kEncThreads) out[i] = 0u;  // slack of the slot
    if (warp == 0) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v |= (uint32_t)nib[8 * lane + k] << (4 * k);
        out[lane] = v;
    } else if (LAY != kLayOne && warp == 1 && lane < kBH / 2) {
        out[32 + lane] = rwc[2 * lane] | (rwc[2 * lane + 1] << 16);
    } else if (warp == 2 && lane < 2) {
        out[hdrw - 2 + lane] = lane ? fl.y : fl.x;
    }
    out += hdrw;
    if (LAY == kLayOne) {
        concat_rows(stage + kStagePad, kStagePitch, rbit0, edges, out, slot - (uint32_t)hdrw, tid);
        return;
    }
    // interleaved region: word k of row r lands at k*bha + r (coalesced stores, conflict-free column reads)
    const uint32_t minw = s_minw, bha = p.bha, inter = minw * bha;
    if (bha == (uint32_t)kBH) {  // thread = (k = warp, r = lane): i = k*32 + r = tid, then k += 8 per step
        const uint32_t *src = stage + lane * kStagePitch + kStagePad + warp;
        uint32_t *dst = out + tid;
        for (uint32_t k = warp; k < minw; k += kEncWarps, src += kEncWarps, dst += kEncThreads) *dst = *src;
    } else {
        for (uint32_t i = tid; i < inter; i += kEncThreads) {
            uint32_t k = i / bha, r = i - k * bha;
            out[i] = stage[r * kStagePitch + kStagePad + k];
        }
    }
    // tails (a few words per row, back to back in row order): eight threads per row
    {
        const uint32_t r = tid >> 3, j = tid & 7;
        if (r < bha) {
            const uint32_t cnt = rwc[r] - minw;
            uint32_t *o = out + inter + (rowoff[r] - r * minw);
            const uint32_t *src = &stage[r * kStagePitch + kStagePad + minw];
            for (uint32_t i = j; i < cnt; i += 8) o[i] = src[i];
        }
    }
}

// status / ticket / ticket_base / epoch: the look-back state of the EXACT layout (shared with the fused encoder); the
// grid is exactly one CTA per block, so the launch consumes n * nb tickets.
void launch_pack(const uint32_t *d_resid, const Geo &g, const uint16_t *d_table, const uint2 *d_flat, uint32_t *d_streams,
                 uint64_t capacity_words, unsigned long long *d_dirE, uint32_t *d_err, unsigned long long *d_status,
                 unsigned long long *d_ticket, unsigned long long ticket_base, uint32_t epoch, cudaStream_t s,
                 const unsigned long long *d_part_base, uint32_t part_hdr_words) {
    uint64_t total = (uint64_t)g.n * g.nb;
    const PackMul pm = {1u << 8, 1u << 10, 1u << 18, 1u << 26};
    const bool g3 = grid3_ok(g) && !(g.flags & FLIC_FLAG_EXACT);
    const dim3 grid = g3 ? dim3(g.nbx, g.nby, g.n) : dim3((unsigned)total);
#define FLIC_PACK2(C, LAY) \
    k_pack<C, LAY><<<grid, kEncThreads, 0, s>>>(d_resid, g, d_table, d_flat, d_streams, capacity_words, d_dirE, d_err, pm, \
                                                d_status, d_ticket, ticket_base, epoch, g3, d_part_base, part_hdr_words)
#define FLIC_PACK(C)                                                          \
    do {                                                                      \
        if (g.flags & FLIC_FLAG_EXACT) FLIC_PACK2(C, kLayExact);              \
        else if (g.flags & FLIC_FLAG_ONE_STREAM) FLIC_PACK2(C, kLayOne);      \
        else FLIC_PACK2(C, kLaySlots);                                        \
    } while (0)
    switch (g.c) {
        case 1: FLIC_PACK(1); break;
        case 2: FLIC_PACK(2); break;
        case 3: FLIC_PACK(3); break;
        default: FLIC_PACK(4); break;
    }
#undef FLIC_PACK
#undef FLIC_PACK2
}

void launch_finalize(const Geo &g, const unsigned long long *d_dirE, uint32_t *d_streams,
                     uint64_t capacity_words, unsigned long long *d_offsets, uint32_t *d_err, cudaStream_t s) {
    uint64_t items = (uint64_t)g.n * (kHdrWords + (uint64_t)g.nb + 1) + g.n + 1;
    uint64_t want = (items + 255) / 256;
    unsigned grid = (unsigned)(want < 148ull * 8 ? want : 148ull * 8);
    k_finalize<<<grid, 256, 0, s>>>(g, d_dirE, d_streams, capacity_words, d_offsets, d_err);
}


// =============================================================================== k_encode (fused)
// Round 2: the whole encode path in ONE pass over the pixels.  A persistent CTA claims blocks in order
// (a ticket counter), and for each block: loads its rows, computes residuals and parks them in a shared
// tile (16 KB that the staged path wrote to and re-read from HBM), histograms them, builds the code table
// with the whole CTA (rank sort, one-thread two-queue merge with both queue heads in registers, depths by
// pointer jumping, depth census, Kraft repair, canonical codes by MATCH.ANY), packs the rows over the same
// tile, obtains its output position with a decoupled look-back over the predecessors' sizes, and copies
// the block out.  DRAM traffic: N read + r.N written (+ 8 B of status and 8 B of directory per block).
//
// The look-back serves all three layouts: with slots (plain FLP0 v3) a block publishes its size as soon
// as its table exists, before packing; with FLIC_FLAG_EXACT the size is only known after packing (the
// pessimistic case: "block sizes are not known before packing"); with FLIC_FLAG_ONE_STREAM the size is
// ceil(code bits / 32) and the rows are concatenated bit-exactly on the way out.
constexpr int kFusedCtas = 6;  // resident CTAs per SM the register budget is set for
constexpr uint32_t kInf = 0x7FFFFFFFu;

struct TabScratch {           // aliases the sub-histograms (dead once every thread holds its symbol's count)
    uint32_t key[256 + 4];    // compacted (count << 8 | symbol), padded with ~0 to a multiple of four
    uint32_t W[256];          // leaf weights, ascending
    uint32_t IW[256];         // internal-node weights in creation order (ascending too)
    uint16_t P[256], D[256];  // parent / distance-to-P of internal nodes (pointer jumping -> depth)
    uint32_t cntd[256];       // internal nodes per depth
    uint8_t ord[256];         // symbol of each sorted rank
    uint8_t lenS[256];        // code length per symbol
    uint32_t num[16];         // leaves per code length
    uint32_t next[16];        // first canonical code per length
    uint32_t wcnt[kEncWarps][16];  // per warp: symbols of each length
    uint32_t wused[kEncWarps];     // per warp: used symbols
    uint32_t bits;
};
static_assert(sizeof(TabScratch) <= kEncWarps * 256 * 4, "table scratch must fit in the sub-histograms");

// FLP0 §3.3 step 2 for one thread: leaves W[0..n) ascending -> parents P[q] of the internal nodes q < n-2
// (node k is created in step k; n-2 is the root).  Only the two queue heads live in registers: the kernel is
// bound by issue slots, not by this loop's latency, so the loop is kept as short as it can be (a version that
// prefetched two entries per queue ran 40 instructions per step against 20 here).
// A leaf wins a tie against an internal node (oracle: flp0_build_lengths).
__device__ __forceinline__ void two_queue_merge(const uint32_t *W, uint32_t *IW, uint16_t *P, int n) {
    int leaf = 0, root = 0;
    uint32_t lw = W[0], iw = kInf;
    for (int k = 0; k < n - 1; ++k) {
        uint32_t sum = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (lw <= iw) {
                sum += lw; ++leaf;
                lw = leaf < n ? W[leaf] : kInf;
            } else {
                sum += iw; P[root] = (uint16_t)k; ++root;
                iw = root < k ? IW[root] : kInf;
            }
        }
        IW[k] = sum;
        if (root == k) iw = sum;  // node k is the only internal node waiting
    }
}

// Code table of one block, built by the whole CTA (256 threads; thread = symbol).  cnt: this symbol's
// count.  Writes tab[s] = code | len << 24 (0: no bits) and nib[s] = length nibble (15: sole symbol);
// returns sum of count x length.  Same rules as oracle/flp0_oracle.c flp0_build_lengths/flp0_assign_codes.
__device__ uint32_t cta_table(uint32_t cnt, TabScratch &t, uint32_t *tab, uint8_t *nib, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t usedm = __ballot_sync(0xFFFFFFFFu, cnt != 0);
    if (lane == 0) t.wused[warp] = __popc(usedm);
    t.lenS[tid] = 0;
    t.cntd[tid] = 0;
    if (tid < 16) t.num[tid] = 0;
    if (tid < kEncWarps * 16) (&t.wcnt[0][0])[tid] = 0;
    if (tid == 0) t.bits = 0;
    __syncthreads();
    int before = 0, n = 0;
#pragma unroll
    for (int w = 0; w < kEncWarps; ++w) { const int u = (int)t.wused[w]; before += w < warp ? u : 0; n += u; }
    if (n <= 1) {  // nothing to code: no symbols at all (every channel flat), or one symbol with a zero-length code
        tab[tid] = 0;
        nib[tid] = cnt ? (uint8_t)kLenSole : (uint8_t)0;
        __syncthreads();
        return 0;
    }
    if (cnt) t.key[before + __popc(usedm & lt)] = (cnt << 8) | (uint32_t)tid;
    if (tid >= n && tid < ((n + 3) & ~3)) t.key[tid] = 0xFFFFFFFFu;
    __syncthreads();
    // (1) sort by (count, symbol): keys are distinct, so a key's rank is the number of smaller keys
    if (tid < n) {
        const uint32_t k = t.key[tid];
        uint32_t r = 0;
        const uint4 *k4 = reinterpret_cast<const uint4 *>(t.key);
        for (int j = 0; j < (n + 3) >> 2; ++j) {
            const uint4 q = k4[j];
            r += (q.x < k) + (q.y < k) + (q.z < k) + (q.w < k);
        }
        t.W[r] = k >> 8;
        t.ord[r] = (uint8_t)k;
    }
    const int root = n - 2;
    if (tid < n - 1) t.D[tid] = tid == root ? 0 : 1;
    if (tid == root) t.P[root] = (uint16_t)root;
    __syncthreads();
    // (2) the merge is serial by nature; everybody else waits at the barrier
    if (tid == 0) two_queue_merge(t.W, t.IW, t.P, n);
    __syncthreads();
    // depth of every internal node: pointer jumping (P <- P[P], D <- D + D[P]) until all point at the root
    for (;;) {
        const uint32_t pq = tid < n - 1 ? t.P[tid] : (uint32_t)root;
        if (!__syncthreads_or(pq != (uint32_t)root)) break;
        const uint32_t d = t.D[tid < n - 1 ? tid : root], pp = t.P[pq], dp = t.D[pq];
        __syncthreads();
        if (tid < n - 1) { t.P[tid] = (uint16_t)pp; t.D[tid] = (uint16_t)(d + dp); }
    }
    // (3) leaves per depth: an internal node at depth d-1 has two children at depth d, internal or leaf;
    // depths beyond kL fold into kL
    if (tid < n - 1) atomicAdd(&t.cntd[t.D[tid]], 1u);
    __syncthreads();
    if (tid >= 1) {
        const uint32_t leaves = 2u * t.cntd[tid - 1] - t.cntd[tid];
        if (leaves) atomicAdd(&t.num[min(tid, kL)], leaves);
    }
    __syncthreads();
    // (4) Kraft repair, then the first canonical code of each length
    if (tid == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int l = 1; l <= kL; ++l) total += t.num[l] << (kL - l);
        while (total > (1u << kL)) {
            t.num[kL]--;
            for (int l = kL - 1; l >= 1; --l)
                if (t.num[l]) { t.num[l]--; t.num[l + 1] += 2; break; }
            --total;
        }
        uint32_t nx = 0, prev = 0;
#pragma unroll
        for (int l = 1; l <= kL; ++l) { nx = (nx + prev) << 1; t.next[l] = nx; prev = t.num[l]; }
    }
    __syncthreads();
    // (5) lengths by sorted rank: the rarest symbols get the longest codes
    if (tid < n) {
        int l = kL;
        uint32_t acc = t.num[kL];
        while ((uint32_t)tid >= acc && l > 1) { --l; acc += t.num[l]; }
        t.lenS[t.ord[tid]] = (uint8_t)l;
    }
    __syncthreads();
    // canonical codes in (length, symbol) order: rank among the equal-length symbols below this one
    const uint32_t l = t.lenS[tid];
    const uint32_t m = __match_any_sync(0xFFFFFFFFu, l);
    if ((m & lt) == 0) t.wcnt[warp][l] = __popc(m);
    __syncthreads();
    uint32_t e = 0;
    if (l) {
        uint32_t r = __popc(m & lt);
#pragma unroll
        for (int w = 0; w < kEncWarps; ++w) r += w < warp ? t.wcnt[w][l] : 0u;
        e = (t.next[l] + r) | (l << 24);
    }
    tab[tid] = e;
    nib[tid] = (uint8_t)l;
    const uint32_t b = __reduce_add_sync(0xFFFFFFFFu, cnt * l);
    if (lane == 0 && b) atomicAdd(&t.bits, b);
    __syncthreads();
    return t.bits;
}

// Stage-level entry (tests): tables from histograms in global memory through cta_table.
__global__ void __launch_bounds__(kEncThreads) k_tables_cta(const uint16_t *__restrict__ hist, uint16_t *__restrict__ table,
                                                          uint32_t *__restrict__ bits) {
    __shared__ __align__(16) TabScratch t;
    __shared__ uint32_t tab[256];
    __shared__ uint8_t nib[256];
    const int tid = threadIdx.x;
    const uint64_t gb = blockIdx.x;
    const uint32_t b = cta_table(hist[gb * 256 + tid], t, tab, nib, tid);
    table[gb * 256 + tid] = (uint16_t)(((uint32_t)nib[tid] << 12) | (tab[tid] & 0xFFFu));
    if (bits && tid == 0) bits[gb] = b;
}

void launch_tables_cta(const uint16_t *d_hist, uint64_t nblocks, uint16_t *d_table, uint32_t *d_bits, cudaStream_t s) {
    k_tables_cta<<<(unsigned)nblocks, kEncThreads, 0, s>>>(d_hist, d_table, d_bits);
}

struct FusedSmem {
    union { uint32_t sh[kEncWarps][256]; TabScratch t; } u;  // first: the sub-histograms must be 1 KB-aligned (hist_inc)
    uint32_t tile[kBH * kStagePitch];  // per row: residual words (word j of lane L at j*32 + L), then the packed row
    uint32_t tile_guard[4];            // ONE_STREAM clears and reads one word past a row: the last row's may be the tile's last
    EdgeWords edges;                   // ONE_STREAM copy-out
    uint32_t tab[256];
    uint8_t nib[256];
    uint32_t rwc[kBH], rowoff[kBH], rbit0[kBH + 1];  // words per row, word offset of each row's tail, first bit of each row
    uint32_t s_or[4], s_first, s_minw, s_used, s_size;
    unsigned long long s_excl;
    long long s_ticket;
};

template <int C, bool SG>
__global__ void __launch_bounds__(kEncThreads, kFusedCtas)
k_encode(const uint8_t *__restrict__ pixels, Geo g, uint32_t *__restrict__ streams, uint64_t capacity_words,
         unsigned long long *__restrict__ dirE, unsigned long long *status, unsigned long long *ticket,
         unsigned long long ticket_base, uint32_t epoch, uint32_t *err, PackMul pm, unsigned long long *phase_clk) {
    __shared__ __align__(1024) FusedSmem sm;
    // phase_clk (debug, normally null): thread 0 of every CTA adds the cycles it spent per phase
    long long t_prev = phase_clk ? clock64() : 0;
#define FLIC_PHASE(i)                                                       \
    if (phase_clk && tid == 0) {                                            \
        const long long t_now = clock64();                                  \
        atomicAdd(phase_clk + (i), (unsigned long long)(t_now - t_prev));   \
        t_prev = t_now;                                                     \
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t total = (uint64_t)g.n * g.nb;
    const bool one = one_stream(g.flags), exact = (g.flags & FLIC_FLAG_EXACT) != 0;
    const int hdrw = one ? kBlkHdrWords1 : kBlkHdrWords;
    const unsigned long long ep = (unsigned long long)epoch << 42;
    const bool fast = g.aligned16 != 0;
    const uint32_t stile = (uint32_t)__cvta_generic_to_shared(sm.tile);

    for (;;) {
        if (tid == 0) sm.s_ticket = (long long)(atomicAdd(ticket, 1ull) - ticket_base);
        {   // the previous block's copy-out has read the tile and the tables: the barrier below orders it
            uint4 *z = reinterpret_cast<uint4 *>(&sm.u.sh[0][0]);
#pragma unroll
            for (int i = 0; i < kEncWarps * 256 / 4 / kEncThreads; ++i) z[tid + i * kEncThreads] = make_uint4(0, 0, 0, 0);
        }
        if (tid < 4) sm.s_or[tid] = 0;
        __syncthreads();
        const uint64_t gb = (uint64_t)sm.s_ticket;
        if (gb >= total) break;
        const BlockPos p = block_pos(g, gb);
        const int nvl = C * max(0, min(4, (int)p.bwa - 4 * lane));  // real bytes of this lane in a real row
        // "all four pixels are real", from its own comparison: with C == 1 ptxas 12.9 turns `nvl == 4` into the
        // predicate output of the VIMNMX that computes the min, and on sm_100a that predicate came out inverted
        // (partial lanes took the full-lane path: caught by the C == 1 ragged-width parity cases)
        const bool lane_full = 4 * lane + 4 <= (int)p.bwa;
        FLIC_PHASE(0)  // ticket + zero-fill + barrier

        // ---- A. rows -> residuals -> tile + sub-histograms -------------------------------------------
        {
            uint32_t v[kBH / kEncWarps][C];
            uint32_t upv[kBH / kEncWarps];
            const uint8_t *row = pixels + (uint64_t)p.img * g.img_stride + (uint64_t)(p.y0 + warp) * g.pitch + (uint64_t)p.x0 * C;
#pragma unroll
            for (int q = 0; q < kBH / kEncWarps; ++q) {  // all loads first: four rows in flight per thread
                const int r = warp + kEncWarps * q;
                int nv;
                upv[q] = 0;
                if (r < (int)p.bha) {
                    load_lane_pixels<C>(row, lane, (int)p.bwa, fast, v[q], &nv);
                    if (lane == 0 && r > 0) upv[q] = up_pixel<C, SG>(row, g.pitch, fast);
                } else {
#pragma unroll
                    for (int j = 0; j < C; ++j) v[q][j] = 0;
                }
                row += kEncWarps * g.pitch;
            }
            uint32_t orw[C];
#pragma unroll
            for (int j = 0; j < C; ++j) orw[j] = 0;
            const uint32_t my = (uint32_t)__cvta_generic_to_shared(sm.u.sh[warp]);
            if (my & 1023u) __trap();  // hist_inc needs 1 KB-aligned sub-histograms (cannot happen: fail loudly)
            uint32_t zero_rows = 0;
#pragma unroll
            for (int q = 0; q < kBH / kEncWarps; ++q) {
                const int r = warp + kEncWarps * q;
                if (r < (int)p.bha) {  // warp-uniform
                    uint32_t res[C];
                    lane_residuals<C, SG>(v[q], upv[q], lane, res);
                    uint32_t *trow = sm.tile + r * kStagePitch + kStagePad + lane;
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        trow[32 * j] = res[j];
                        uint32_t x = res[j];
                        if (q == 0 && j == 0 && tid == 0) {  // the block's first pixel: its residual is its value
                            constexpr uint32_t fm = C == 4 ? 0xFFFFFFFFu : ((1u << (8 * (C & 3))) - 1u);
                            sm.s_first = x & fm;
                            x &= ~fm;
                        }
                        // bytes past the lane's last real pixel hold garbage differences: keep them out of the OR
                        const int nb = min(4, max(0, nvl - 4 * j));
                        const uint32_t vm = nb == 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
                        orw[C == 4 ? 0 : j] |= x & vm;
                    }
                    if (C == 4 && __all_sync(0xFFFFFFFFu, lane_full && ((res[0] | res[1 % C] | res[2 % C] | res[3 % C]) & 0xFF000000u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 4 * C; ++j)
                            if ((j & 3) != 3) hist_inc(my, res[j >> 2], j & 3);
                        ++zero_rows;  // an all-zero alpha row (opaque plane): counted once, not looked up
                    } else if (lane_full) {
#pragma unroll
                        for (int j = 0; j < 4 * C; ++j) hist_inc(my, res[j >> 2], j & 3);
                    } else if (nvl > 0) {
#pragma unroll
                        for (int j = 0; j < 4 * C; ++j)
                            if (j < nvl) hist_inc(my, res[j >> 2], j & 3);
                    }
                }
            }
            if (C == 4 && lane == 0 && zero_rows) atomicAdd(&sm.u.sh[warp][0], zero_rows * (uint32_t)kBW);
            constexpr int NA = C == 4 ? 1 : C;
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                const uint32_t o = __reduce_or_sync(0xFFFFFFFFu, orw[j]);
                if (lane == 0 && o) atomicOr(&sm.s_or[j], o);
            }
        }
        __syncthreads();
        FLIC_PHASE(1)  // loads, residuals, sub-histograms

        // ---- B. this thread's symbol count; flat channels (FLP0 §2b) ----------------------------------
        uint32_t cnt = 0, flatmask, flatvals;
        {
#pragma unroll
            for (int k = 0; k < kEncWarps; ++k) cnt += sm.u.sh[k][tid];
            uint32_t T;
            if (C == 4) T = sm.s_or[0];
            else if (C == 3) {  // word j byte b carries channel (4j + b) mod 3
                const uint32_t a = sm.s_or[0], b = sm.s_or[1 % C], c = sm.s_or[2 % C];
                T = (a | __byte_perm(b, 0u, 0x4102) | __byte_perm(c, 0u, 0x4021) | (a >> 24) |
                     __byte_perm(b, 0u, 0x4434) | __byte_perm(c, 0u, 0x4344)) & 0x00FFFFFFu;
            } else if (C == 2) { const uint32_t x = sm.s_or[0] | sm.s_or[1 % C]; T = (x | (x >> 16)) & 0xFFFFu; }
            else { uint32_t x = sm.s_or[0]; x |= x >> 16; T = (x | (x >> 8)) & 0xFFu; }
            auto nonzero_bytes = [](uint32_t x) { x |= x >> 4; x |= x >> 2; x |= x >> 1; return x & 0x01010101u; };
            constexpr uint32_t chb = C == 4 ? 0x01010101u : ((1u << (8 * (C & 3))) - 1u) & 0x01010101u;
            const uint32_t flatb = ~nonzero_bytes(T) & chb;  // bit 8*ch: channel ch is flat
            const uint32_t first = sm.s_first, npix = p.bwa * p.bha;
            if (tid == 0) cnt -= __popc(flatb) * (npix - 1u);
            cnt -= __popc(~nonzero_bytes(first ^ ((uint32_t)tid * 0x01010101u)) & flatb);
            flatmask = (flatb * 0x01020408u) >> 24;
            flatvals = first & (flatb * 0xFFu);
        }
        __syncthreads();  // every thread has read the sub-histograms: the table scratch may overwrite them
        FLIC_PHASE(2)  // histogram reduce, flat channels

        // ---- C. code table ------------------------------------------------------------------------------
        const uint32_t code_bits = cta_table(cnt, sm.u.t, sm.tab, sm.nib, tid);
        uint32_t size = 0;  // words this block occupies in the stream (known now, except with EXACT)
        if (one) size = (uint32_t)kBlkHdrWords1 + ((code_bits + 31u) >> 5);
        else if (!exact) size = (uint32_t)kBlkHdrWords + (code_bits >> 5) + (code_bits ? p.bha : 0u);
        if (!exact && tid == 0 && gb > 0) st_status(status + gb, ep | kStA | size);  // early: successors need not wait for the packing

        FLIC_PHASE(3)  // code table

        // ---- D. pack the rows over the tile --------------------------------------------------------------
        {
            uint32_t skip[C];
#pragma unroll
            for (int j = 0; j < C; ++j) skip[j] = word_channel_bits<C>(flatmask, j);
#pragma unroll 1
            for (int q = 0; q < kBH / kEncWarps; ++q) {
                const int r = warp + kEncWarps * q;
                uint32_t *trow = sm.tile + r * kStagePitch + kStagePad;
                const int nv = r < (int)p.bha ? nvl : 0;
                const bool full = lane_full && r < (int)p.bha;
                uint32_t cur[C];
#pragma unroll
                for (int j = 0; j < C; ++j) cur[j] = nv ? trow[32 * j + lane] : 0u;
                uint32_t qlo[C], qhi[C], ql[C], nbits = 0;
                if (full && flatmask == 0) {
#pragma unroll
                    for (int j = 0; j < C; ++j) { ql[j] = quad_of<true, 0>(cur[j], 4 * j, 4 * C, 0u, sm.tab, pm, qlo[j], qhi[j]); nbits += ql[j]; }
                } else if (full && C == 4 && flatmask == 8u) {
#pragma unroll
                    for (int j = 0; j < C; ++j) { ql[j] = quad_of<true, 8>(cur[j], 4 * j, 4 * C, 8u, sm.tab, pm, qlo[j], qhi[j]); nbits += ql[j]; }
                } else {
#pragma unroll
                    for (int j = 0; j < C; ++j) { ql[j] = quad_of<false, -1>(cur[j], 4 * j, nv, skip[j], sm.tab, pm, qlo[j], qhi[j]); nbits += ql[j]; }
                }
                const uint32_t incl = warp_incl_scan(nbits, lane);
                const uint32_t rowbits = __shfl_sync(0xFFFFFFFFu, incl, 31);
                const uint32_t rw = (rowbits + 31u) >> 5;
                if (lane == 31) { sm.rwc[r] = rw; sm.rbit0[r] = rowbits; }
                __syncwarp();  // all lanes hold their residuals: the row may be overwritten
                // clear exactly the words the copy-out will read (one more with ONE_STREAM, whose funnel
                // shifts look one word ahead)
                for (uint32_t i = lane; i < rw + (one ? 1u : 0u); i += 32) trow[i] = 0u;
                __syncwarp();
                uint32_t nb = 0u - (8u * (stile + 4u * (uint32_t)(r * kStagePitch + kStagePad)) + (incl - nbits));
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    nb -= ql[j];
                    const uint32_t a = ((31u - nb) >> 3) & ~3u;  // byte address one past the word the group ends in
                    red_or_shared(a, -4, __funnelshift_l(0u, qlo[j], nb));
                    red_or_shared(a, -8, __funnelshift_l(qlo[j], qhi[j], nb));
                    if (__any_sync(0xFFFFFFFFu, ql[j] > 32u)) red_or_shared(a, -12, __funnelshift_l(qhi[j], 0u, nb));
                }
            }
        }
        __syncthreads();
        FLIC_PHASE(4)  // pack

        // ---- E. size, position (decoupled look-back), directory ---------------------------------------------
        if (warp == 0) {
            const uint32_t wcount = sm.rwc[lane];
            const uint32_t incl = warp_incl_scan(wcount, lane);
            sm.rowoff[lane] = incl - wcount;
            uint32_t mn = lane < (int)p.bha ? wcount : 0xFFFFFFFFu;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
            const uint32_t rb = sm.rbit0[lane];
            const uint32_t bincl = warp_incl_scan(rb, lane);
            __syncwarp();
            sm.rbit0[lane] = bincl - rb;
            if (lane == 31) sm.rbit0[kBH] = bincl;
            uint32_t used = one ? (uint32_t)kBlkHdrWords1 + ((__shfl_sync(0xFFFFFFFFu, bincl, 31) + 31u) >> 5)
                                : (uint32_t)kBlkHdrWords + __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (exact) size = used;
            if (used > size || (one && used != size)) {  // cannot happen: a slot covers the rows' padding; one stream has none
                if (lane == 0) atomicOr(err, kErrSlot);
                used = 0xFFFFFFFFu;
            }
            // exclusive prefix of the sizes of all blocks before this one
            unsigned long long excl = 0;
            if (gb > 0) {
                if (exact && lane == 0) st_status(status + gb, ep | kStA | size);
                excl = lookback_excl(status, gb, epoch, lane);
            }
            if (lane == 0) {
                st_status(status + gb, ep | kStP | ((excl + size) & kStVal));
                dirE[gb] = excl;
                if (gb + 1 == total) dirE[total] = excl + size;
                sm.s_excl = excl;
                sm.s_minw = mn;
                sm.s_used = used;
                sm.s_size = size;
            }
        }
        __syncthreads();
        FLIC_PHASE(5)  // look-back
        size = sm.s_size;
        const uint32_t used = sm.s_used;
        const unsigned long long base = (unsigned long long)(p.img + 1) * (kHdrWords + g.nb + 1) + sm.s_excl;
        if (used == 0xFFFFFFFFu) continue;
        if (base + size > capacity_words) {
            if (tid == 0) atomicOr(err, kErrCapacity);
            continue;
        }

        // ---- F. copy-out --------------------------------------------------------------------------------------
        uint32_t *out = streams + base;
        if (warp == 0) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) v |= (uint32_t)sm.nib[8 * lane + k] << (4 * k);
            out[lane] = v;
        } else if (!one && warp == 1 && lane < kBH / 2) {
            out[32 + lane] = sm.rwc[2 * lane] | (sm.rwc[2 * lane + 1] << 16);
        } else if (warp == 2 && lane < 2) {
            out[hdrw - 2 + lane] = lane ? flatvals : flatmask;
        }
        out += hdrw;
        if (one) {
            concat_rows(sm.tile + kStagePad, kStagePitch, sm.rbit0, sm.edges, out, size - (uint32_t)kBlkHdrWords1, tid);
            continue;
        }
        for (uint32_t i = used + tid; i < size; i += kEncThreads) (out - hdrw)[i] = 0u;  // slack of the slot
        const uint32_t minw = sm.s_minw, bha = p.bha, inter = minw * bha;
        if (bha == (uint32_t)kBH) {  // thread = (k = warp, r = lane): i = k*32 + r = tid, then k += 8 per step
            const uint32_t *src = sm.tile + lane * kStagePitch + kStagePad + warp;
            uint32_t *dst = out + tid;
            for (uint32_t k = warp; k < minw; k += kEncWarps, src += kEncWarps, dst += kEncThreads) *dst = *src;
        } else {
            for (uint32_t i = tid; i < inter; i += kEncThreads) {
                const uint32_t k = i / bha, r = i - k * bha;
                out[i] = sm.tile[r * kStagePitch + kStagePad + k];
            }
        }
        {   // tails (a few words per row, back to back in row order): eight threads per row
            const uint32_t r = tid >> 3, j = tid & 7;
            if (r < bha) {
                const uint32_t cntw = sm.rwc[r] - minw;
                uint32_t *o = out + inter + (sm.rowoff[r] - r * minw);
                const uint32_t *src = &sm.tile[r * kStagePitch + kStagePad + minw];
                for (uint32_t i = j; i < cntw; i += 8) o[i] = src[i];
            }
        }
        FLIC_PHASE(6)  // copy-out (this thread's share)
    }
#undef FLIC_PHASE
}

// Grid: resident CTAs only (the loop is persistent); correctness does not depend on co-residency, because a
// CTA that has not started has not claimed a ticket, and tickets are what the look-back waits on.
template <int C, bool SG>
static unsigned fused_grid(uint64_t total) {
    static int per_sm = 0, sms = 0;
    if (!per_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode<C, SG>, kEncThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const uint64_t cap = (uint64_t)sms * per_sm;
    return (unsigned)(total < cap ? total : cap);
}

unsigned launch_encode_fused(const uint8_t *d_pixels, const Geo &g, uint32_t *d_streams, uint64_t capacity_words,
                             unsigned long long *d_dirE, unsigned long long *d_status, unsigned long long *d_ticket,
                             unsigned long long ticket_base, uint32_t epoch, uint32_t *d_err, unsigned long long *d_phase_clk,
                             cudaStream_t s) {
    const uint64_t total = (uint64_t)g.n * g.nb;
    const bool sg = (g.flags & FLIC_FLAG_SUBGREEN) && g.c >= 3;
    const PackMul pm = {1u << 8, 1u << 10, 1u << 18, 1u << 26};
    unsigned grid = 0;
#define FLIC_ENC(C, SG) \
    (grid = fused_grid<C, SG>(total), \
     k_encode<C, SG><<<grid, kEncThreads, 0, s>>>(d_pixels, g, d_streams, capacity_words, d_dirE, d_status, d_ticket, ticket_base, epoch, d_err, pm, d_phase_clk))
    switch (g.c) {
        case 1: FLIC_ENC(1, false); break;
        case 2: FLIC_ENC(2, false); break;
        case 3: if (sg) FLIC_ENC(3, true); else FLIC_ENC(3, false); break;
        default: if (sg) FLIC_ENC(4, true); else FLIC_ENC(4, false); break;
    }
#undef FLIC_ENC
    return grid;
}

}  // namespace flic
