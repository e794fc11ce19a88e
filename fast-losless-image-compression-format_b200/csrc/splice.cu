// splice.cu — device side of the block-row splice ("one oversized image split by block rows"):
// FLP0 blocks never predict across block edges, so the streams of runs of whole block rows concatenate;
// only the directory needs rebasing and the header rewriting.  The payloads and directory entries are
// put in place by copies (D2D on one GPU, NCCL send/recv straight into the output across GPUs); these
// kernels finish the job.  Format: DESIGN.md §FLP0 (provisional; not the reference's bitstream).
#include "common.cuh"

namespace flic {

__device__ __forceinline__ void write_header(uint32_t *out, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, uint32_t nb,
                                             uint32_t pw, int k) {
    uint32_t v;
    switch (k) {
        case 0: v = kMagic; break;
        case 1: v = kVersion | (c << 16) | ((flags & 0xFFu) << 24); break;
        case 2: v = w; break;
        case 3: v = h; break;
        case 4: v = (uint32_t)kBW | ((uint32_t)kBH << 16); break;
        case 5: v = nb; break;
        case 6: v = pw; break;
        default: v = (uint32_t)kL; break;
    }
    out[k] = v;
}

// out: [8 header words][nb + 1 directory words][payload]; the directory holds each part's own (part-relative)
// offsets, part j's entries at first_block[j] ..: add the payload words of the parts before it.
__global__ void __launch_bounds__(256) k_splice_finish(uint32_t *out, SpliceParts sp, uint32_t w, uint32_t h, uint32_t c,
                                                       uint32_t flags) {
    const uint32_t nb = sp.first_block[sp.k], pw = sp.base_words[sp.k];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    if (tid < kHdrWords) write_header(out, w, h, c, flags, nb, pw, (int)tid);
    if (tid == kHdrWords) out[kHdrWords + nb] = pw;
    for (uint64_t i = tid; i < nb; i += stride) {
        uint32_t j = 0;  // the part block i belongs to (k <= 64: a short scan of a kernel-parameter array)
        while (j + 1 < sp.k && sp.first_block[j + 1] <= i) ++j;
        out[kHdrWords + i] += sp.base_words[j];
    }
}

// part: [8 words to be written][nb + 1 directory words copied verbatim from the full stream][payload].
// Two launches: the rebase reads entry 0 as the base, so entry 0 itself is rewritten afterwards.
__global__ void __launch_bounds__(256) k_split_rebase(uint32_t *part, uint32_t nb) {
    const uint32_t first = part[kHdrWords];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = 1 + tid; i <= nb; i += stride) part[kHdrWords + i] -= first;
}
__global__ void k_split_head(uint32_t *part, uint32_t nb, uint32_t w, uint32_t h, uint32_t c, uint32_t flags) {
    const int t = threadIdx.x;
    if (t < kHdrWords) write_header(part, w, h, c, flags, nb, part[kHdrWords + nb], t);
    if (t == kHdrWords) part[kHdrWords] = 0;
}

// ---- block-row split with peer memory: every GPU writes its part straight into the spliced stream (k_pack with a part
// base) and its directory entries with k_part_directory; one GPU adds the header; for decode a GPU pulls its part out.
// dirE: the part's own exclusive slot prefix (k_slots); dir_out: the spliced directory at this part's first block.
__global__ void __launch_bounds__(256) k_part_directory(const unsigned long long *__restrict__ dirE, uint32_t nb,
                                                        const unsigned long long *__restrict__ base_words, uint32_t *dir_out) {
    const unsigned long long base = *base_words;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = tid; i < nb; i += stride) dir_out[i] = (uint32_t)(base + dirE[i]);
}
__global__ void k_part_words(const unsigned long long *__restrict__ dirE, uint32_t nb, unsigned long long *out) {
    if (threadIdx.x == 0) *out = dirE[nb];
}
__global__ void k_splice_header(uint32_t *out, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, uint32_t nb,
                                const unsigned long long *__restrict__ total_words, uint64_t capacity_words, uint32_t *err) {
    const int t = threadIdx.x;
    const unsigned long long pw = *total_words;
    if (pw > 0xFFFFFFFFull) { if (t == 0) atomicOr(err, kErrRange); return; }
    if ((unsigned long long)kHdrWords + nb + 1 + pw > capacity_words) { if (t == 0) atomicOr(err, kErrCapacity); return; }
    if (t < kHdrWords) write_header(out, w, h, c, flags, nb, (uint32_t)pw, t);
    if (t == kHdrWords) out[kHdrWords + nb] = (uint32_t)pw;
}
// part <- [8 words for the header][directory entries first_block .. first_block + nb][the payload words they span], read
// from a stream that may live in another GPU's memory (coalesced 4-byte accesses: source and destination are only
// word-aligned relative to each other).  part_bytes receives the size of the part stream once finished.
__global__ void __launch_bounds__(256) k_pull_part(const uint32_t *__restrict__ stream, uint64_t stream_words, uint32_t total_blocks,
                                                   uint32_t first_block, uint32_t nb, uint32_t *__restrict__ part,
                                                   uint64_t capacity_words, unsigned long long *part_bytes, uint32_t *err) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    // the stream is input: nothing is read past what its (validated) header says it holds, nothing past its buffer.
    // One thread per CTA reads the five words (they may be a NVLink round trip away) and shares the verdict.
    const uint64_t fixed = (uint64_t)kHdrWords + total_blocks + 1;
    __shared__ uint32_t s_b0, s_b1, s_state;  // state: 0 ok, else the error bit
    if (threadIdx.x == 0) {
        bool ok = stream_words >= fixed && stream[0] == kMagic && stream[5] == total_blocks && fixed + stream[6] <= stream_words;
        uint32_t b0 = 0, b1 = 0;
        if (ok) {
            b0 = stream[kHdrWords + first_block];
            b1 = stream[kHdrWords + first_block + nb];
            ok = b0 <= b1 && b1 <= stream[6];
        }
        const bool fits = ok && (uint64_t)kHdrWords + nb + 1 + (b1 - b0) <= capacity_words;
        s_b0 = b0; s_b1 = b1; s_state = fits ? 0u : (ok ? kErrCapacity : kErrFormat);
    }
    __syncthreads();
    const uint32_t b0 = s_b0, b1 = s_b1;
    if (s_state) {
        if (tid == 0) { atomicOr(err, s_state); *part_bytes = 0; }
        return;
    }
    if (tid == 0) *part_bytes = 4ull * ((uint64_t)kHdrWords + nb + 1 + (b1 - b0));
    const uint32_t *dir = stream + kHdrWords + first_block;
    for (uint64_t i = tid; i <= nb; i += stride) part[kHdrWords + i] = dir[i];
    // payload: 16-byte loads on the (possibly remote) source side — a NVLink round trip is ~2 us, so what is in flight per
    // thread decides the rate (4-byte loads from 300 K threads are 1.2 MB in flight: not enough for 900 GB/s) — and word
    // stores on the local side, since source and destination are only word-aligned relative to each other
    const uint32_t *src = stream + fixed + b0;
    uint32_t *dst = part + kHdrWords + nb + 1;
    const uint64_t n = b1 - b0;
    const uint64_t head = min(n, (uint64_t)((4u - (uint32_t)(((uintptr_t)src >> 2) & 3u)) & 3u));
    const uint64_t n4 = (n - head) >> 2, tail0 = head + 4 * n4;
    if (tid < head) dst[tid] = src[tid];
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
    for (uint64_t i = tid; i < n4; i += stride) {
        const uint4 v = s4[i];
        uint32_t *d = dst + head + 4 * i;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    if (tid < n - tail0) dst[tail0 + tid] = src[tail0 + tid];
}

void launch_part_words(const unsigned long long *d_dirE, uint32_t part_blocks, unsigned long long *d_part_words_out, cudaStream_t s) {
    k_part_words<<<1, 32, 0, s>>>(d_dirE, part_blocks, d_part_words_out);
}
void launch_part_directory(const unsigned long long *d_dirE, uint32_t part_blocks, const unsigned long long *d_base_words,
                           uint32_t *d_dir_out, cudaStream_t s) {
    const unsigned grid = (unsigned)((part_blocks + 255) / 256 < 1 ? 1 : ((part_blocks + 255) / 256 > 1184 ? 1184 : (part_blocks + 255) / 256));
    k_part_directory<<<grid, 256, 0, s>>>(d_dirE, part_blocks, d_base_words, d_dir_out);
}
void launch_splice_header(uint32_t *d_out, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, uint32_t nb,
                          const unsigned long long *d_total_words, uint64_t capacity_words, uint32_t *d_err, cudaStream_t s) {
    k_splice_header<<<1, 32, 0, s>>>(d_out, w, h, c, flags, nb, d_total_words, capacity_words, d_err);
}
void launch_pull_part(const uint32_t *d_stream, uint64_t stream_words, uint32_t total_blocks, uint32_t first_block, uint32_t part_blocks,
                      uint32_t *d_part, uint64_t capacity_words, unsigned long long *d_part_bytes, uint32_t *d_err, cudaStream_t s) {
    k_pull_part<<<148 * 8, 256, 0, s>>>(d_stream, stream_words, total_blocks, first_block, part_blocks, d_part, capacity_words, d_part_bytes,
                                        d_err);
}

void launch_splice_finish(uint32_t *d_out, const SpliceParts &sp, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                          cudaStream_t s) {
    const uint32_t nb = sp.first_block[sp.k];
    const unsigned grid = (unsigned)((nb + 255) / 256 < 1 ? 1 : ((nb + 255) / 256 > 1184 ? 1184 : (nb + 255) / 256));
    k_splice_finish<<<grid, 256, 0, s>>>(d_out, sp, w, h, c, flags);
}

void launch_split_finish(uint32_t *d_part, uint32_t nb, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, cudaStream_t s) {
    const unsigned grid = (unsigned)((nb + 255) / 256 < 1 ? 1 : ((nb + 255) / 256 > 1184 ? 1184 : (nb + 255) / 256));
    k_split_rebase<<<grid, 256, 0, s>>>(d_part, nb);
    k_split_head<<<1, 32, 0, s>>>(d_part, nb, w, h, c, flags);
}

}  // namespace flic
