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
