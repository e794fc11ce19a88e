// api.cu — C ABI of libflicb200.so (include/flic_b200.h): context, workspace,
// device-resident and host-buffer batch entry points, header parsing and the
// block-row splice.  Host logic only; every byte of codec work happens in the
// kernels of encode.cu / decode.cu.  There is no CPU fallback anywhere here.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cuda.h>

#include "common.cuh"

using namespace flic;

struct flic_ctx {
    int device = 0;
    char msg[256] = {0};
    uint64_t launches = 0;
    // per-block workspace (grown on demand)
    uint64_t ws_blocks = 0;
    uint16_t *d_hist = nullptr, *d_table = nullptr;
    uint32_t *d_resid = nullptr;  // residual plane: 32 rows x 32 lanes x C words (<= 16 KB) per block
    uint2 *d_flat = nullptr;      // per block {flat-channel mask, values}
    uint32_t *d_bits = nullptr;             // per block: sum of count x code length
    unsigned long long *d_dirE = nullptr;   // per block: exclusive prefix sum of slot words (+ grand total)
    unsigned long long *d_slot_status = nullptr;  // k_slots: 128 epoch-tagged run sums
    uint32_t slot_epoch = 0;
    uint32_t *d_err = nullptr;
    uint32_t *h_err = nullptr;  // pinned
    // host-API pipeline: chunks of the batch flow H2D -> kernels -> D2H on three streams with
    // double-buffered device staging, so PCIe in, compute and PCIe out overlap (grown on demand)
    uint8_t *d_pix[2] = {nullptr, nullptr}, *d_str[2] = {nullptr, nullptr};
    unsigned long long *d_off[2] = {nullptr, nullptr};
    unsigned long long *h_off[2] = {nullptr, nullptr};  // pinned, off_cap entries each
    uint64_t pix_cap = 0, str_cap = 0, off_cap = 0;
    cudaStream_t stream = nullptr;                      // kernels of the host-buffer API
    cudaStream_t s_in = nullptr, s_out = nullptr;       // H2D / D2H
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // opt-in per-kernel timing (flic_set_kernel_timing): event pairs recorded on the launching stream
    bool timing = false;
    struct Span { cudaEvent_t a, b; int kernel; };
    std::vector<Span> spans;      // recorded, not yet read
    std::vector<Span> free_spans; // recycled events
};

namespace {
// Brackets one kernel launch with events when timing is on; a no-op otherwise.
struct KernelTimer {
    flic_ctx *ctx; cudaStream_t s; flic_ctx::Span sp; bool on;
    KernelTimer(flic_ctx *c, int kernel, cudaStream_t st) : ctx(c), s(st), on(c->timing) {
        if (!on) return;
        if (!ctx->free_spans.empty()) { sp = ctx->free_spans.back(); ctx->free_spans.pop_back(); }
        else if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) { on = false; return; }
        sp.kernel = kernel;
        cudaEventRecord(sp.a, s);
    }
    ~KernelTimer() {
        if (!on) return;
        cudaEventRecord(sp.b, s);
        ctx->spans.push_back(sp);
    }
};
}  // namespace

static int cuda_fail(flic_ctx *ctx, cudaError_t e, const char *what) {
    if (ctx) snprintf(ctx->msg, sizeof ctx->msg, "%s: %s", what, cudaGetErrorString(e));
    return FLIC_E_CUDA;
}
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);   \
    } while (0)

static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

extern "C" int flic_version(void) { return (int)kVersion; }

extern "C" const char *flic_strerror(int code) {
    switch (code) {
        case FLIC_OK: return "ok";
        case FLIC_E_ARG: return "invalid argument";
        case FLIC_E_CAPACITY: return "output buffer too small";
        case FLIC_E_FORMAT: return "malformed stream";
        case FLIC_E_CUDA: return "CUDA error";
        case FLIC_E_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
        case FLIC_E_UNSUPPORTED: return "unsupported format feature";
        case FLIC_E_INTERNAL: return "device-side consistency check failed";
        default: return "unknown error";
    }
}

extern "C" const char *flic_last_error(const flic_ctx *ctx) { return ctx ? ctx->msg : ""; }
extern "C" uint64_t flic_launch_count(const flic_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" uint64_t flic_blocks_per_image(uint32_t w, uint32_t h) {
    return (uint64_t)cdiv(w, kBW) * cdiv(h, kBH);
}

extern "C" uint64_t flic_max_stream_bytes(uint32_t w, uint32_t h, uint32_t c) {
    uint64_t nb = flic_blocks_per_image(w, h);
    uint64_t blk = kBlkHdrWords + (uint64_t)kBH * ((kBW * c * kL + 31) / 32 + 1);  // + slot slack: one word per row
    return 4ull * (kHdrWords + nb + 1 + nb * blk);
}

extern "C" int flic_create(int device, flic_ctx **out) {
    if (!out) return FLIC_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return FLIC_E_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return FLIC_E_NO_DEVICE;
    flic_ctx *ctx = new (std::nothrow) flic_ctx;
    if (!ctx) return FLIC_E_ARG;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_err, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_err, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_slot_status, 128 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_slot_status, 0, 128 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_err, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        flic_destroy(ctx);
        return FLIC_E_CUDA;
    }
    *out = ctx;
    return FLIC_OK;
}

extern "C" void flic_destroy(flic_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_hist); cudaFree(ctx->d_table); cudaFree(ctx->d_bits); cudaFree(ctx->d_dirE);
    cudaFree(ctx->d_resid); cudaFree(ctx->d_flat); cudaFree(ctx->d_err); cudaFree(ctx->d_slot_status);
    for (int i = 0; i < 2; ++i) {
        cudaFree(ctx->d_pix[i]); cudaFree(ctx->d_str[i]); cudaFree(ctx->d_off[i]);
        if (ctx->h_off[i]) cudaFreeHost(ctx->h_off[i]);
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
        if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    for (auto &sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto &sp : ctx->free_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    delete ctx;
}

static int ensure_workspace(flic_ctx *ctx, uint64_t blocks) {
    if (blocks <= ctx->ws_blocks) return FLIC_OK;
    cudaFree(ctx->d_hist); cudaFree(ctx->d_table); cudaFree(ctx->d_bits); cudaFree(ctx->d_dirE);
    cudaFree(ctx->d_resid); cudaFree(ctx->d_flat);
    ctx->d_hist = ctx->d_table = nullptr; ctx->d_bits = nullptr; ctx->d_dirE = nullptr; ctx->d_resid = nullptr; ctx->d_flat = nullptr;
    ctx->ws_blocks = 0;
    CU(cudaMalloc(&ctx->d_resid, blocks * (uint64_t)kBH * 512));
    CU(cudaMalloc(&ctx->d_flat, blocks * sizeof(uint2)));
    CU(cudaMalloc(&ctx->d_hist, blocks * 256 * sizeof(uint16_t)));
    CU(cudaMalloc(&ctx->d_table, blocks * 256 * sizeof(uint16_t)));
    CU(cudaMalloc(&ctx->d_bits, blocks * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_dirE, (blocks + 1) * sizeof(unsigned long long)));
    ctx->ws_blocks = blocks;
    return FLIC_OK;
}

static int make_geo(const void *base, uint32_t n, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, Geo *g) {
    if (n == 0 || w == 0 || h == 0 || c < 1 || c > 4) return FLIC_E_ARG;
    if ((flags & 0x0Fu) != FLIC_PRED_LEFT || (flags & ~0x1Fu)) return FLIC_E_ARG;
    g->n = n; g->w = w; g->h = h; g->c = c; g->flags = flags;
    g->nbx = cdiv(w, kBW); g->nby = cdiv(h, kBH); g->nb = g->nbx * g->nby;
    g->pitch = (uint64_t)w * c;
    g->img_stride = g->pitch * h;
    g->aligned16 = (((uintptr_t)base | g->pitch | g->img_stride) & 15u) == 0;
    g->aligned32 = (((uintptr_t)base | g->pitch | g->img_stride) & 31u) == 0;
    if ((uint64_t)n * g->nb >= (1ull << 31)) return FLIC_E_ARG;
    return FLIC_OK;
}

extern "C" int flic_stage_histograms(flic_ctx *ctx, const uint8_t *d_pixels, uint32_t n, uint32_t w, uint32_t h,
                                     uint32_t c, uint32_t flags, uint16_t *d_hist, uint32_t *d_flat, void *stream) {
    if (!ctx || !d_pixels || !d_hist) return FLIC_E_ARG;
    Geo g;
    int rc = make_geo(d_pixels, n, w, h, c, flags, &g);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    { KernelTimer t(ctx, FLIC_K_HISTOGRAMS, (cudaStream_t)stream); launch_histograms(d_pixels, g, d_hist, nullptr, reinterpret_cast<uint2 *>(d_flat), (cudaStream_t)stream); }
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_stage_tables(flic_ctx *ctx, const uint16_t *d_hist, uint64_t n_blocks_total, uint16_t *d_table,
                                 uint32_t *d_bits, void *stream) {
    if (!ctx || !d_hist || !d_table || n_blocks_total == 0) return FLIC_E_ARG;
    CU(cudaSetDevice(ctx->device));
    { KernelTimer t(ctx, FLIC_K_TABLES, (cudaStream_t)stream); launch_tables(d_hist, n_blocks_total, d_table, d_bits, (cudaStream_t)stream); }
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_encode_batch_device(flic_ctx *ctx, const uint8_t *d_pixels, uint32_t n, uint32_t w, uint32_t h,
                                        uint32_t c, uint32_t flags, uint8_t *d_streams, uint64_t capacity_bytes,
                                        uint64_t *d_offsets, void *stream) {
    if (!ctx || !d_pixels || !d_streams || !d_offsets || ((uintptr_t)d_streams & 15u)) return FLIC_E_ARG;
    Geo g;
    int rc = make_geo(d_pixels, n, w, h, c, flags, &g);
    if (rc) return rc;
    if (capacity_bytes < 4ull * n * (kHdrWords + (uint64_t)g.nb + 1)) return FLIC_E_CAPACITY;
    CU(cudaSetDevice(ctx->device));
    rc = ensure_workspace(ctx, (uint64_t)n * g.nb);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t cap_words = capacity_bytes / 4;
    { KernelTimer t(ctx, FLIC_K_HISTOGRAMS, s); launch_histograms(d_pixels, g, ctx->d_hist, ctx->d_resid, ctx->d_flat, s); }
    { KernelTimer t(ctx, FLIC_K_TABLES, s); launch_tables(ctx->d_hist, (uint64_t)n * g.nb, ctx->d_table, ctx->d_bits, s); }
    bool fused;
    { KernelTimer t(ctx, FLIC_K_SLOTS, s);
      fused = launch_slots(g, ctx->d_bits, ctx->d_dirE, ctx->d_slot_status, ++ctx->slot_epoch, cap_words, ctx->d_err,
                           (uint32_t *)d_streams, (unsigned long long *)d_offsets, s); }
    if (!fused) {  // headers and directories depend on the slots only: before k_pack, off its tail
        KernelTimer t(ctx, FLIC_K_FINALIZE, s);
        launch_finalize(g, ctx->d_dirE, (uint32_t *)d_streams, cap_words, (unsigned long long *)d_offsets, ctx->d_err, s);
    }
    { KernelTimer t(ctx, FLIC_K_PACK, s);
      launch_pack(ctx->d_resid, g, ctx->d_table, ctx->d_flat, (uint32_t *)d_streams, cap_words, ctx->d_dirE, ctx->d_err, s); }
    ctx->launches += fused ? 4 : 5;
    CU(cudaGetLastError());
    return FLIC_OK;
}

// TMA descriptor of a tightly packed RGBA pixel batch for k_decode's store path: 3-D {row bytes, rows,
// images}, 64 B x 32 rows boxes, 64 B swizzle.  Returns false when the layout does not qualify (then the
// kernel stores directly) or the driver entry point is missing.
static bool make_pixel_map(const Geo &g, uint8_t *d_pixels, CUtensorMap *tm) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (encode_fn)fn;
    }();
    static const bool off = getenv("FLIC_NO_TMA") != nullptr;
    if (off || !encode || g.c != 4 || !g.aligned16) return false;
    const cuuint64_t dims[3] = {g.pitch, g.h, g.n};
    const cuuint64_t strides[2] = {g.pitch, g.img_stride};
    const cuuint32_t box[3] = {64, (cuuint32_t)kBH, 1}, estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_pixels, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int flic_decode_batch_device(flic_ctx *ctx, const uint8_t *d_streams, const uint64_t *d_offsets, uint32_t n,
                                        uint32_t w, uint32_t h, uint32_t c, uint32_t flags, uint8_t *d_pixels,
                                        void *stream) {
    if (!ctx || !d_pixels || !d_streams || !d_offsets || ((uintptr_t)d_streams & 3u)) return FLIC_E_ARG;
    Geo g;
    int rc = make_geo(d_pixels, n, w, h, c, flags, &g);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    { KernelTimer t(ctx, FLIC_K_DECODE, (cudaStream_t)stream);
      alignas(64) CUtensorMap tm;
      const bool tma = make_pixel_map(g, d_pixels, &tm);
      launch_decode((const uint32_t *)d_streams, (const unsigned long long *)d_offsets, g, d_pixels, ctx->d_err,
                    tma ? &tm : nullptr, (cudaStream_t)stream); }
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_set_kernel_timing(flic_ctx *ctx, int enable) {
    if (!ctx) return FLIC_E_ARG;
    ctx->timing = enable != 0;
    return FLIC_OK;
}

extern "C" int flic_get_kernel_times(flic_ctx *ctx, double ms[FLIC_K_COUNT], uint64_t counts[FLIC_K_COUNT]) {
    if (!ctx || !ms || !counts) return FLIC_E_ARG;
    for (int i = 0; i < FLIC_K_COUNT; ++i) { ms[i] = 0.0; counts[i] = 0; }
    for (auto &sp : ctx->spans) {
        CU(cudaEventSynchronize(sp.b));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, sp.a, sp.b));
        ms[sp.kernel] += t;
        counts[sp.kernel] += 1;
        ctx->free_spans.push_back(sp);
    }
    ctx->spans.clear();
    return FLIC_OK;
}

extern "C" int flic_check(flic_ctx *ctx, void *stream) {
    if (!ctx) return FLIC_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(uint32_t), s));
    CU(cudaStreamSynchronize(s));
    uint32_t e = *ctx->h_err;
    if (!e) return FLIC_OK;
    snprintf(ctx->msg, sizeof ctx->msg, "device error bits 0x%x%s%s%s", e, (e & kErrCapacity) ? " capacity" : "",
             (e & kErrSlot) ? " slot-overrun" : "", (e & kErrFormat) ? " format" : "");
    if (e & kErrFormat) return FLIC_E_FORMAT;
    if (e & kErrCapacity) return FLIC_E_CAPACITY;
    return FLIC_E_INTERNAL;
}

// ------------------------------------------------------------ host-buffer API
static int ensure_staging(flic_ctx *ctx, uint64_t pix, uint64_t str, uint64_t noff) {
    if (pix > ctx->pix_cap) {
        for (int i = 0; i < 2; ++i) { cudaFree(ctx->d_pix[i]); ctx->d_pix[i] = nullptr; }
        ctx->pix_cap = 0;
        for (int i = 0; i < 2; ++i) CU(cudaMalloc(&ctx->d_pix[i], pix));
        ctx->pix_cap = pix;
    }
    if (str > ctx->str_cap) {
        for (int i = 0; i < 2; ++i) { cudaFree(ctx->d_str[i]); ctx->d_str[i] = nullptr; }
        ctx->str_cap = 0;
        for (int i = 0; i < 2; ++i) CU(cudaMalloc(&ctx->d_str[i], str + 16));
        ctx->str_cap = str;
    }
    if (noff > ctx->off_cap) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(ctx->d_off[i]); ctx->d_off[i] = nullptr;
            if (ctx->h_off[i]) cudaFreeHost(ctx->h_off[i]);
            ctx->h_off[i] = nullptr;
        }
        ctx->off_cap = 0;
        for (int i = 0; i < 2; ++i) {
            CU(cudaMalloc(&ctx->d_off[i], noff * sizeof(unsigned long long)));
            CU(cudaMallocHost(&ctx->h_off[i], noff * sizeof(unsigned long long)));
        }
        ctx->off_cap = noff;
    }
    return FLIC_OK;
}

// images per pipeline chunk: ~64 MB of pixels.  The call is PCIe-bound, so what a chunk size buys is a short
// pipeline fill and drain (measured on 64 x 4K RGBA: 256 MB chunks 23.3 GB/s, 64 MB 24.5, 33 MB 24.6).
static uint32_t chunk_images(uint32_t n, uint64_t image_bytes) {
    uint64_t target = 64ull << 20;
    if (const char *e = getenv("FLIC_CHUNK_BYTES")) {  // test hook: force many small chunks
        unsigned long long v = strtoull(e, nullptr, 10);
        if (v) target = v;
    }
    uint64_t m = target / (image_bytes ? image_bytes : 1);
    if (m < 1) m = 1;
    return (uint32_t)(m > n ? n : m);
}

static int drain(flic_ctx *ctx) {  // after a failure: leave no work in flight on the staging buffers
    cudaStreamSynchronize(ctx->s_in); cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->s_out);
    return 0;
}

extern "C" int flic_encode_batch(flic_ctx *ctx, const uint8_t *h_pixels, uint32_t n, uint32_t w, uint32_t h, uint32_t c,
                                 uint32_t flags, uint8_t *h_streams, uint64_t capacity_bytes, uint64_t *h_offsets) {
    if (!ctx || !h_pixels || !h_streams || !h_offsets) return FLIC_E_ARG;
    if (n == 0 || w == 0 || h == 0 || c < 1 || c > 4) return FLIC_E_ARG;
    if ((flags & 0x0Fu) != FLIC_PRED_LEFT || (flags & ~0x1Fu)) return FLIC_E_ARG;
    const uint64_t img_bytes = (uint64_t)w * h * c, img_worst = flic_max_stream_bytes(w, h, c);
    const uint32_t m = chunk_images(n, img_bytes);
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_staging(ctx, m * img_bytes, m * img_worst, (uint64_t)m + 1);
    if (rc) return rc;
    const uint32_t chunks = (n + m - 1) / m;
    uint64_t out_pos = 0;
    h_offsets[0] = 0;
    // software pipeline: H2D(k+1) is issued before the host waits for the sizes of chunk k
    auto issue_in = [&](uint32_t k) -> int {
        const int b = k & 1;
        const uint32_t first = k * m, cnt = (first + m <= n) ? m : n - first;
        if (k >= 2) CU(cudaStreamWaitEvent(ctx->s_in, ctx->ev_k[b], 0));  // kernels of chunk k-2 have consumed d_pix[b]
        CU(cudaMemcpyAsync(ctx->d_pix[b], h_pixels + (uint64_t)first * img_bytes, cnt * img_bytes, cudaMemcpyHostToDevice,
                           ctx->s_in));
        CU(cudaEventRecord(ctx->ev_in[b], ctx->s_in));
        return FLIC_OK;
    };
    rc = issue_in(0);
    for (uint32_t k = 0; k < chunks && rc == FLIC_OK; ++k) {
        const int b = k & 1;
        const uint32_t first = k * m, cnt = (first + m <= n) ? m : n - first;
        rc = [&]() -> int {
            CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
            if (k >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));  // D2H of chunk k-2 has drained d_str[b]
            int r = flic_encode_batch_device(ctx, ctx->d_pix[b], cnt, w, h, c, flags, ctx->d_str[b], cnt * img_worst,
                                             (uint64_t *)ctx->d_off[b], ctx->stream);
            if (r) return r;
            CU(cudaMemcpyAsync(ctx->h_off[b], ctx->d_off[b], ((uint64_t)cnt + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaEventRecord(ctx->ev_k[b], ctx->stream));
            if (k + 1 < chunks) { r = issue_in(k + 1); if (r) return r; }
            CU(cudaEventSynchronize(ctx->ev_k[b]));
            const uint64_t total = ctx->h_off[b][cnt];
            if (total > cnt * img_worst) return FLIC_E_INTERNAL;  // kernels flagged a capacity overrun
            if (out_pos + total > capacity_bytes) return FLIC_E_CAPACITY;
            for (uint32_t i = 1; i <= cnt; ++i) h_offsets[first + i] = out_pos + ctx->h_off[b][i];
            CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[b], 0));
            CU(cudaMemcpyAsync(h_streams + out_pos, ctx->d_str[b], total, cudaMemcpyDeviceToHost, ctx->s_out));
            CU(cudaEventRecord(ctx->ev_out[b], ctx->s_out));
            out_pos += total;
            return FLIC_OK;
        }();
    }
    if (rc) { drain(ctx); flic_check(ctx, ctx->stream); return rc; }
    CU(cudaStreamSynchronize(ctx->s_out));
    return flic_check(ctx, ctx->stream);
}

extern "C" int flic_peek(const uint8_t *s, uint64_t size, flic_image_info *info) {
    if (!s || !info || size < FLIC_HEADER_BYTES) return FLIC_E_FORMAT;
    uint32_t wd[8];
    memcpy(wd, s, sizeof wd);
    if (wd[0] != kMagic || (wd[1] & 0xFFFFu) != kVersion || wd[7] != (uint32_t)kL) return FLIC_E_FORMAT;
    info->channels = (wd[1] >> 16) & 0xFFu;
    info->flags = wd[1] >> 24;
    info->width = wd[2];
    info->height = wd[3];
    info->block_w = wd[4] & 0xFFFFu;
    info->block_h = wd[4] >> 16;
    info->n_blocks = wd[5];
    info->payload_words = wd[6];
    if (info->width == 0 || info->height == 0 || info->channels < 1 || info->channels > 4) return FLIC_E_FORMAT;
    if ((info->flags & 0x0Fu) != FLIC_PRED_LEFT || (info->flags & ~0x1Fu)) return FLIC_E_FORMAT;
    if (info->block_w == 0 || info->block_h == 0 || (info->block_h & 1)) return FLIC_E_FORMAT;
    if ((uint64_t)cdiv(info->width, info->block_w) * cdiv(info->height, info->block_h) != info->n_blocks)
        return FLIC_E_FORMAT;
    if (4ull * (kHdrWords + (uint64_t)info->n_blocks + 1 + info->payload_words) > size) return FLIC_E_FORMAT;
    return FLIC_OK;
}

extern "C" int flic_decode_batch(flic_ctx *ctx, const uint8_t *h_streams, const uint64_t *h_offsets, uint32_t n,
                                 uint8_t *h_pixels, uint64_t pixels_capacity) {
    if (!ctx || !h_streams || !h_offsets || !h_pixels || n == 0) return FLIC_E_ARG;
    flic_image_info first;
    for (uint32_t i = 0; i < n; ++i) {
        if (h_offsets[i + 1] < h_offsets[i] || (h_offsets[i] & 3u)) return FLIC_E_FORMAT;
        flic_image_info info;
        int rc = flic_peek(h_streams + h_offsets[i], h_offsets[i + 1] - h_offsets[i], &info);
        if (rc) return rc;
        if (info.block_w != (uint32_t)kBW || info.block_h != (uint32_t)kBH) return FLIC_E_UNSUPPORTED;
        if (i == 0) first = info;
        else if (info.width != first.width || info.height != first.height || info.channels != first.channels ||
                 info.flags != first.flags)
            return FLIC_E_UNSUPPORTED;  // one launch decodes one geometry; split mixed batches by geometry
    }
    const uint64_t img_bytes = (uint64_t)first.width * first.height * first.channels;
    if ((uint64_t)n * img_bytes > pixels_capacity) return FLIC_E_CAPACITY;
    const uint32_t m = chunk_images(n, img_bytes);
    const uint32_t chunks = (n + m - 1) / m;
    uint64_t max_str = 0;
    for (uint32_t k = 0; k < chunks; ++k) {
        const uint32_t f0 = k * m, f1 = (f0 + m <= n) ? f0 + m : n;
        const uint64_t sz = h_offsets[f1] - h_offsets[f0];
        if (sz > max_str) max_str = sz;
    }
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_staging(ctx, m * img_bytes, max_str, (uint64_t)m + 1);
    if (rc) return rc;
    for (uint32_t k = 0; k < chunks; ++k) {
        const int b = k & 1;
        const uint32_t f0 = k * m, cnt = (f0 + m <= n) ? m : n - f0;
        const uint64_t base = h_offsets[f0], sz = h_offsets[f0 + cnt] - base;
        // the pinned offsets and d_str[b] of chunk k-2 must have been consumed by its kernel
        if (k >= 2) { CU(cudaEventSynchronize(ctx->ev_k[b])); }
        for (uint32_t i = 0; i <= cnt; ++i) ctx->h_off[b][i] = h_offsets[f0 + i] - base;
        CU(cudaMemcpyAsync(ctx->d_str[b], h_streams + base, sz, cudaMemcpyHostToDevice, ctx->s_in));
        CU(cudaMemcpyAsync(ctx->d_off[b], ctx->h_off[b], ((uint64_t)cnt + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
        CU(cudaEventRecord(ctx->ev_in[b], ctx->s_in));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
        if (k >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));  // D2H of chunk k-2 has drained d_pix[b]
        rc = flic_decode_batch_device(ctx, ctx->d_str[b], (const uint64_t *)ctx->d_off[b], cnt, first.width, first.height,
                                      first.channels, first.flags, ctx->d_pix[b], ctx->stream);
        if (rc) { drain(ctx); return rc; }
        CU(cudaEventRecord(ctx->ev_k[b], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[b], 0));
        CU(cudaMemcpyAsync(h_pixels + (uint64_t)f0 * img_bytes, ctx->d_pix[b], cnt * img_bytes, cudaMemcpyDeviceToHost,
                           ctx->s_out));
        CU(cudaEventRecord(ctx->ev_out[b], ctx->s_out));
    }
    CU(cudaStreamSynchronize(ctx->s_out));
    return flic_check(ctx, ctx->stream);
}

// ------------------------------------------------------------------- splice
extern "C" int flic_splice_block_rows(const uint8_t *const *parts, const uint64_t *part_sizes, uint32_t k, uint8_t *out,
                                      uint64_t out_capacity, uint64_t *out_size) {
    if (!parts || !part_sizes || !out || !out_size || k == 0) return FLIC_E_ARG;
    flic_image_info first, info;
    uint64_t nb = 0, pw = 0, height = 0;
    for (uint32_t i = 0; i < k; ++i) {
        int rc = flic_peek(parts[i], part_sizes[i], &info);
        if (rc) return rc;
        if (i == 0) first = info;
        else if (info.width != first.width || info.channels != first.channels || info.flags != first.flags ||
                 info.block_w != first.block_w || info.block_h != first.block_h)
            return FLIC_E_ARG;
        if (i + 1 < k && info.height % info.block_h) return FLIC_E_ARG;  // only the last part may be ragged
        nb += info.n_blocks; pw += info.payload_words; height += info.height;
    }
    if (nb >= (1ull << 32) || pw >= (1ull << 32) || height >= (1ull << 32)) return FLIC_E_ARG;
    const uint64_t total = 4ull * (kHdrWords + nb + 1 + pw);
    if (total > out_capacity) return FLIC_E_CAPACITY;
    uint32_t hdr[8] = {kMagic, kVersion | (first.channels << 16) | (first.flags << 24), first.width, (uint32_t)height,
                       first.block_w | (first.block_h << 16), (uint32_t)nb, (uint32_t)pw, (uint32_t)kL};
    memcpy(out, hdr, sizeof hdr);
    uint8_t *dir = out + 4 * kHdrWords, *payload = dir + 4 * (nb + 1);
    uint32_t base = 0;
    for (uint32_t i = 0; i < k; ++i) {
        flic_peek(parts[i], part_sizes[i], &info);
        const uint8_t *pdir = parts[i] + 4 * kHdrWords;
        for (uint32_t b = 0; b < info.n_blocks; ++b) {
            uint32_t v;
            memcpy(&v, pdir + 4ull * b, 4);
            v += base;
            memcpy(dir, &v, 4);
            dir += 4;
        }
        memcpy(payload, pdir + 4ull * (info.n_blocks + 1), 4ull * info.payload_words);
        payload += 4ull * info.payload_words;
        base += info.payload_words;
    }
    memcpy(dir, &base, 4);
    *out_size = total;
    return FLIC_OK;
}
