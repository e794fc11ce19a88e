// api.cu — C ABI of libflicb200.so (include/flic_b200.h): context, workspace,
// device-resident and host-buffer batch entry points (blocking and submit/wait),
// header parsing and the block-row splice (host and device).  Host logic only; every
// byte of codec work happens in the kernels of encode.cu / decode.cu / decode_one.cu.
// There is no CPU fallback anywhere here.
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <cuda.h>

#include "common.cuh"

using namespace flic;

namespace {
// Staging of one direction of the host-buffer API: chunks of the batch flow H2D -> kernels -> D2H on three
// streams with double-buffered device buffers, so PCIe in, compute and PCIe out overlap.  Encode and decode
// own separate pipes, so an encode call and a decode call (flic_*_submit) can be in flight together and use
// both directions of the link at once.
// Up to kDepth chunks are in flight per call.  How many actually are is decided chunk by chunk (measured on a B200
// box, 64 x 4K RGBA: tools/e2e_probe.py, profiles/r02_link_probe.txt):
//   * a call that has the link to itself wants a deep pipeline: 42 ms at depth >= 3 against 57 ms at depth 1;
//   * an encode and a decode in flight TOGETHER want depth 1: 68.5 ms for both (the bare link does the same copy
//     pattern in 67.6 ms), against 86 ms at any depth >= 2 — exactly the sum of the two calls alone.  With more than
//     one copy per call queued, the copy engines serve the two calls' same-direction copies in convoys and the
//     opposite direction idles.
// So the lookahead is kDepth while the other direction's pipe is idle and 1 while it is active.  Two more rules, from
// the same measurements: a copy is issued only when it can start (the host has already seen the event it depends
// on), and the encoder's stream sizes come back through mapped pinned memory the kernels write to, not through a
// device-to-host copy queued behind the kernels.
constexpr int kDepth = 4;
// deepest lookahead a call may use (FLIC_PIPE_DEPTH: experiment switch, 1..kDepth)
static int pipe_depth() {
    static const int d = [] { const char *e = getenv("FLIC_PIPE_DEPTH"); const int v = e ? atoi(e) : kDepth; return v < 1 ? 1 : (v > kDepth ? kDepth : v); }();
    return d;
}
// lookahead while the other direction's call is in flight (FLIC_PIPE_BOTH_DEPTH: experiment switch)
static int pipe_depth_both() {
    static const int d = [] { const char *e = getenv("FLIC_PIPE_BOTH_DEPTH"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : (v > kDepth ? kDepth : v); }();
    return d;
}
// Host wait for a chunk's kernels.  At depth 1 the wait is on the critical path of every chunk, so it polls (a blocking
// wait costs a sleep/wake-up per chunk); with a deep pipeline it blocks, leaving the cores to the other ranks' workers.
static cudaError_t wait_chunk(cudaEvent_t ev, bool poll) {
    static const int mode = [] { const char *e = getenv("FLIC_WAIT"); return e ? (strcmp(e, "poll") == 0 ? 1 : (strcmp(e, "block") == 0 ? 2 : 0)) : 0; }();
    if (mode == 1) poll = true;
    if (mode == 2) poll = false;
    if (!poll) return cudaEventSynchronize(ev);
    for (;;) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) return e;
    }
}
struct Pipe {
    uint8_t *d_pix[kDepth] = {}, *d_str[kDepth] = {};
    unsigned long long *d_off[kDepth] = {};
    unsigned long long *h_off[kDepth] = {};  // pinned + mapped, off_cap entries each
    unsigned long long *m_off[kDepth] = {};  // the device's view of h_off (encode kernels write their offsets straight to the host)
    uint64_t pix_cap = 0, str_cap = 0, off_cap = 0;
    cudaStream_t s_k = nullptr, s_in = nullptr, s_out = nullptr;  // kernels / H2D / D2H
    cudaEvent_t ev_in[kDepth] = {}, ev_k[kDepth] = {}, ev_out[kDepth] = {};
    std::atomic<int> active{0};  // a call is running through this pipe (the other pipe throttles its lookahead)
    // submit/wait
    std::thread worker;
    bool busy = false;
    int result = FLIC_OK;
};
}  // namespace

struct flic_ctx {
    int device = 0;
    char msg[256] = {0};
    std::atomic<uint64_t> launches{0};  // encode and decode workers (flic_*_submit) may both be launching
    int encoder = FLIC_ENCODER_AUTO;
    // per-block workspace (grown on demand)
    uint64_t ws_blocks = 0;
    bool ws_staged = false;       // the staged encoder's extra arrays are allocated
    uint16_t *d_hist = nullptr, *d_table = nullptr;
    uint32_t *d_resid = nullptr;  // staged encoder: residual plane, 32 rows x 32 lanes x C words (<= 16 KB) per block
    uint2 *d_flat = nullptr;      // staged encoder: per block {flat-channel mask, values}
    uint32_t *d_bits = nullptr;             // staged encoder: per block sum of count x code length
    unsigned long long *d_dirE = nullptr;   // per block: exclusive prefix sum of block sizes in words (+ grand total)
    unsigned long long *d_status = nullptr; // fused encoder: per block look-back status (epoch-tagged, never memset)
    unsigned long long *d_ticket = nullptr; // fused encoder: block ticket counter (monotonic across launches)
    unsigned long long ticket_base = 0;
    uint32_t fused_epoch = 0;
    unsigned long long *d_phase = nullptr;  // debug: per-phase cycle sums of k_encode (FLIC_PHASE_CLOCKS=1), else null
    unsigned long long *d_slot_status = nullptr;  // staged encoder, k_slots: 128 epoch-tagged run sums
    uint32_t slot_epoch = 0;
    int slots_max_grid = 0;       // co-resident k_slots CTAs on this device (occupancy API)
    // the workspace is one per context: an encode on another stream first waits for the previous one
    cudaEvent_t ev_ws = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;
    // device-side error flags: word 0 is raised by encode kernels, word 1 by decode kernels, so that an encode
    // and a decode in flight together (flic_*_submit) each report their own.  h_err: four pinned words —
    // [0] / [1] are read by the encode / decode pipelines, [2..3] by flic_check
    uint32_t *d_err = nullptr;
    Geo plan_geo{};               // flic_encode_plan_device -> flic_encode_emit_device
    bool plan_valid = false;
    uint32_t *h_err = nullptr;
    std::mutex span_mu;
    Pipe enc, dec;
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
        bool reuse;
        {
            std::lock_guard<std::mutex> lk(ctx->span_mu);
            reuse = !ctx->free_spans.empty();
            if (reuse) { sp = ctx->free_spans.back(); ctx->free_spans.pop_back(); }
        }
        if (!reuse && (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess)) { on = false; return; }
        sp.kernel = kernel;
        cudaEventRecord(sp.a, s);
    }
    ~KernelTimer() {
        if (!on) return;
        cudaEventRecord(sp.b, s);
        std::lock_guard<std::mutex> lk(ctx->span_mu);
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

// overflow-safe ceil(a / b)
static inline uint64_t cdiv(uint64_t a, uint64_t b) { return a / b + (a % b != 0); }

static inline bool flags_ok(uint32_t flags) {
    return (flags & 0x0Fu) == FLIC_PRED_LEFT && (flags & ~FLIC_FLAGS_ALL) == 0 &&
           !((flags & FLIC_FLAG_ONE_STREAM) && (flags & FLIC_FLAG_EXACT));
}

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
        case FLIC_E_BUSY: return "an operation submitted on this context has not been waited for";
        default: return "unknown error";
    }
}

extern "C" const char *flic_last_error(const flic_ctx *ctx) { return ctx ? ctx->msg : ""; }
extern "C" uint64_t flic_launch_count(const flic_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

extern "C" uint64_t flic_blocks_per_image(uint32_t w, uint32_t h) { return cdiv(w, kBW) * cdiv(h, kBH); }

extern "C" uint64_t flic_max_stream_bytes(uint32_t w, uint32_t h, uint32_t c) {
    uint64_t nb = flic_blocks_per_image(w, h);
    uint64_t blk = kBlkHdrWords + (uint64_t)kBH * ((kBW * c * kL + 31) / 32 + 1);  // + slot slack: one word per row
    return 4ull * (kHdrWords + nb + 1 + nb * blk);
}

static cudaError_t pipe_create(Pipe &p) {
    cudaError_t e = cudaStreamCreateWithFlags(&p.s_k, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking);
    for (int i = 0; i < kDepth && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p.ev_in[i], cudaEventDisableTiming);
        // the host waits on ev_k once per chunk: a blocking (not spinning) wait, so that the other call's worker thread
        // is not starved of the driver while this one waits
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.ev_k[i], cudaEventDisableTiming | cudaEventBlockingSync);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.ev_out[i], cudaEventDisableTiming);
    }
    return e;
}

static void pipe_destroy(Pipe &p) {
    if (p.worker.joinable()) p.worker.join();
    for (int i = 0; i < kDepth; ++i) {
        cudaFree(p.d_pix[i]); cudaFree(p.d_str[i]); cudaFree(p.d_off[i]);
        if (p.h_off[i]) cudaFreeHost(p.h_off[i]);
        if (p.ev_in[i]) cudaEventDestroy(p.ev_in[i]);
        if (p.ev_k[i]) cudaEventDestroy(p.ev_k[i]);
        if (p.ev_out[i]) cudaEventDestroy(p.ev_out[i]);
    }
    if (p.s_k) cudaStreamDestroy(p.s_k);
    if (p.s_in) cudaStreamDestroy(p.s_in);
    if (p.s_out) cudaStreamDestroy(p.s_out);
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
    if (const char *e = getenv("FLIC_ENCODER"))
        ctx->encoder = strcmp(e, "staged") == 0 ? FLIC_ENCODER_STAGED : (strcmp(e, "fused") == 0 ? FLIC_ENCODER_FUSED : FLIC_ENCODER_AUTO);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_err, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_err, 0, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_slot_status, 128 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_slot_status, 0, 128 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_ticket, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_ticket, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_err, 4 * sizeof(uint32_t));
    if (e == cudaSuccess && getenv("FLIC_PHASE_CLOCKS")) {
        e = cudaMalloc(&ctx->d_phase, 2 * FLIC_PHASES * sizeof(unsigned long long));  // k_encode, then k_decode_one
        if (e == cudaSuccess) e = cudaMemset(ctx->d_phase, 0, 2 * FLIC_PHASES * sizeof(unsigned long long));
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_ws, cudaEventDisableTiming);
    if (e == cudaSuccess) e = pipe_create(ctx->enc);
    if (e == cudaSuccess) e = pipe_create(ctx->dec);
    if (e == cudaSuccess) ctx->slots_max_grid = slots_max_resident_ctas();
    if (e != cudaSuccess || ctx->slots_max_grid < 1) {
        flic_destroy(ctx);
        return FLIC_E_CUDA;
    }
    *out = ctx;
    return FLIC_OK;
}

static void free_workspace(flic_ctx *ctx) {
    cudaFree(ctx->d_hist); cudaFree(ctx->d_table); cudaFree(ctx->d_bits); cudaFree(ctx->d_dirE);
    cudaFree(ctx->d_resid); cudaFree(ctx->d_flat); cudaFree(ctx->d_status);
    ctx->d_hist = ctx->d_table = nullptr; ctx->d_bits = nullptr; ctx->d_dirE = nullptr; ctx->d_resid = nullptr;
    ctx->d_flat = nullptr; ctx->d_status = nullptr;
    ctx->ws_blocks = 0; ctx->ws_staged = false;
}

extern "C" void flic_destroy(flic_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    pipe_destroy(ctx->enc);
    pipe_destroy(ctx->dec);
    cudaDeviceSynchronize();
    free_workspace(ctx);
    cudaFree(ctx->d_err); cudaFree(ctx->d_slot_status); cudaFree(ctx->d_ticket); cudaFree(ctx->d_phase);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->ev_ws) cudaEventDestroy(ctx->ev_ws);
    for (auto &sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto &sp : ctx->free_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    delete ctx;
}

extern "C" int flic_set_option(flic_ctx *ctx, int option, int value) {
    if (!ctx) return FLIC_E_ARG;
    switch (option) {
        case FLIC_OPT_ENCODER:
            if (value != FLIC_ENCODER_FUSED && value != FLIC_ENCODER_STAGED && value != FLIC_ENCODER_AUTO) return FLIC_E_ARG;
            ctx->encoder = value;
            return FLIC_OK;
        default: return FLIC_E_ARG;
    }
}

static int ensure_workspace(flic_ctx *ctx, uint64_t blocks, bool staged) {
    if (blocks <= ctx->ws_blocks && (!staged || ctx->ws_staged)) return FLIC_OK;
    CU(cudaDeviceSynchronize());  // nothing may still be using the arrays about to be freed
    if (blocks < ctx->ws_blocks) blocks = ctx->ws_blocks;
    staged = staged || ctx->ws_staged;
    free_workspace(ctx);
    CU(cudaMalloc(&ctx->d_dirE, (blocks + 1) * sizeof(unsigned long long)));
    CU(cudaMalloc(&ctx->d_status, blocks * sizeof(unsigned long long)));
    CU(cudaMemset(ctx->d_status, 0, blocks * sizeof(unsigned long long)));
    ctx->fused_epoch = 0;
    if (staged) {
        CU(cudaMalloc(&ctx->d_resid, blocks * (uint64_t)kBH * 512));
        CU(cudaMalloc(&ctx->d_flat, blocks * sizeof(uint2)));
        CU(cudaMalloc(&ctx->d_hist, blocks * 256 * sizeof(uint16_t)));
        CU(cudaMalloc(&ctx->d_table, blocks * 256 * sizeof(uint16_t)));
        CU(cudaMalloc(&ctx->d_bits, blocks * sizeof(uint32_t)));
    }
    ctx->ws_blocks = blocks;
    ctx->ws_staged = staged;
    return FLIC_OK;
}

static int make_geo(const void *base, uint32_t n, uint32_t w, uint32_t h, uint32_t c, uint32_t flags, Geo *g) {
    if (n == 0 || w == 0 || h == 0 || c < 1 || c > 4) return FLIC_E_ARG;
    if (!flags_ok(flags)) return FLIC_E_ARG;
    const uint64_t nbx = cdiv(w, kBW), nby = cdiv(h, kBH), nb = nbx * nby;
    if (nb > 0xFFFFFFFFull || (uint64_t)n * nb >= (1ull << 31)) return FLIC_E_ARG;  // kernels index blocks in 31 bits
    g->n = n; g->w = w; g->h = h; g->c = c; g->flags = flags;
    g->nbx = (uint32_t)nbx; g->nby = (uint32_t)nby; g->nb = (uint32_t)nb;
    g->pitch = (uint64_t)w * c;
    g->img_stride = g->pitch * h;
    g->aligned16 = (((uintptr_t)base | g->pitch | g->img_stride) & 15u) == 0;
    g->aligned32 = (((uintptr_t)base | g->pitch | g->img_stride) & 31u) == 0;
    return FLIC_OK;
}

static bool make_tile_map(const Geo &g, const uint8_t *d_pixels, CUtensorMapDataType dt, uint32_t elem_bytes, uint32_t box_elems,
                          CUtensorMapSwizzle swz, CUtensorMap *tm);
// TMA descriptor for the encoder's pixel tile loads: u32 elements, a box of one whole block (32 rows x 128*C bytes),
// no swizzle (lanes read 4*C consecutive bytes: conflict-free as it is).  FLIC_NO_TMA=1 turns both TMA paths off.
static bool make_load_map(const Geo &g, const uint8_t *d_pixels, CUtensorMap *tm) {
    static const bool off = getenv("FLIC_NO_TMA_LOAD") != nullptr;  // A/B switch for the load path alone
    if (off) return false;
    return make_tile_map(g, d_pixels, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 32u * g.c, CU_TENSOR_MAP_SWIZZLE_NONE, tm);
}

extern "C" int flic_stage_histograms(flic_ctx *ctx, const uint8_t *d_pixels, uint32_t n, uint32_t w, uint32_t h,
                                     uint32_t c, uint32_t flags, uint16_t *d_hist, uint32_t *d_flat, void *stream) {
    if (!ctx || !d_pixels || !d_hist) return FLIC_E_ARG;
    Geo g;
    int rc = make_geo(d_pixels, n, w, h, c, flags, &g);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    { KernelTimer t(ctx, FLIC_K_HISTOGRAMS, (cudaStream_t)stream);
      alignas(64) CUtensorMap tm;
      const bool tma = make_load_map(g, d_pixels, &tm);
      launch_histograms(d_pixels, g, d_hist, nullptr, reinterpret_cast<uint2 *>(d_flat), tma ? &tm : nullptr, (cudaStream_t)stream); }
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_stage_tables(flic_ctx *ctx, const uint16_t *d_hist, uint64_t n_blocks_total, uint16_t *d_table,
                                 uint32_t *d_bits, void *stream) {
    if (!ctx || !d_hist || !d_table || n_blocks_total == 0 || n_blocks_total >= (1ull << 31)) return FLIC_E_ARG;
    CU(cudaSetDevice(ctx->device));
    { KernelTimer t(ctx, FLIC_K_TABLES, (cudaStream_t)stream);
      if (ctx->encoder != FLIC_ENCODER_STAGED) launch_tables_cta(d_hist, n_blocks_total, d_table, d_bits, (cudaStream_t)stream);  // the fused kernel's builder
      else launch_tables(d_hist, n_blocks_total, d_table, d_bits, (cudaStream_t)stream); }
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
    // AUTO: below kAutoBlocks the job is a latency chain (launches, one wave of CTAs) and the two-launch fused path
    // wins; above it the staged pipeline's higher issue efficiency does (measured: DESIGN.md §5)
    constexpr uint64_t kAutoBlocks = 4096;
    const bool staged = ctx->encoder == FLIC_ENCODER_STAGED || (ctx->encoder == FLIC_ENCODER_AUTO && (uint64_t)n * g.nb >= kAutoBlocks);
    const bool exact = (flags & FLIC_FLAG_EXACT) != 0;
    CU(cudaSetDevice(ctx->device));
    rc = ensure_workspace(ctx, (uint64_t)n * g.nb, staged);
    if (rc) return rc;
    ctx->plan_valid = false;  // this call reuses the workspace a pending flic_encode_plan_device left its plan in
    cudaStream_t s = (cudaStream_t)stream;
    if (ctx->ws_used && ctx->ws_stream != s) CU(cudaStreamWaitEvent(s, ctx->ev_ws, 0));  // one workspace per context
    const uint64_t cap_words = capacity_bytes / 4;
    if (!staged || exact) {  // a launch that looks back: a fresh epoch on the status array
        if (++ctx->fused_epoch >= (1u << 22)) {  // the 22-bit epoch wrapped: start over on a clean status array
            CU(cudaMemsetAsync(ctx->d_status, 0, ctx->ws_blocks * sizeof(unsigned long long), s));
            ctx->fused_epoch = 1;
        }
    }
    if (!staged) {
        // fused single pass (k_encode) + headers/directories (k_finalize)
        unsigned grid;
        { KernelTimer t(ctx, FLIC_K_ENCODE, s);
          grid = launch_encode_fused(d_pixels, g, (uint32_t *)d_streams, cap_words, ctx->d_dirE, ctx->d_status, ctx->d_ticket,
                                     ctx->ticket_base, ctx->fused_epoch, ctx->d_err, ctx->d_phase, s); }
        ctx->ticket_base += (uint64_t)n * g.nb + grid;  // every CTA's last claim fails
        { KernelTimer t(ctx, FLIC_K_FINALIZE, s);
          launch_finalize(g, ctx->d_dirE, (uint32_t *)d_streams, cap_words, (unsigned long long *)d_offsets, ctx->d_err, s); }
        ctx->launches += 2;
    } else {
        { KernelTimer t(ctx, FLIC_K_HISTOGRAMS, s);
          alignas(64) CUtensorMap tm;
          const bool tma = make_load_map(g, d_pixels, &tm);
          launch_histograms(d_pixels, g, ctx->d_hist, ctx->d_resid, ctx->d_flat, tma ? &tm : nullptr, s); }
        { KernelTimer t(ctx, FLIC_K_TABLES, s); launch_tables(ctx->d_hist, (uint64_t)n * g.nb, ctx->d_table, ctx->d_bits, s); }
        // Slots and ONE_STREAM: every block's position follows from its histogram and code lengths (k_slots), and the
        // headers and directories from the positions — before k_pack, off its tail.  EXACT: positions exist only once the
        // blocks are packed (k_pack looks back over the packed sizes), so k_finalize comes last and there is no k_slots.
        bool fused = false;
        if (!exact) {
            { KernelTimer t(ctx, FLIC_K_SLOTS, s);
              fused = launch_slots(g, ctx->d_bits, ctx->d_dirE, ctx->d_slot_status, ++ctx->slot_epoch, cap_words, ctx->d_err,
                                   (uint32_t *)d_streams, (unsigned long long *)d_offsets, ctx->slots_max_grid, s); }
            if (!fused) {
                KernelTimer t(ctx, FLIC_K_FINALIZE, s);
                launch_finalize(g, ctx->d_dirE, (uint32_t *)d_streams, cap_words, (unsigned long long *)d_offsets, ctx->d_err, s);
            }
        }
        { KernelTimer t(ctx, FLIC_K_PACK, s);
          launch_pack(ctx->d_resid, g, ctx->d_table, ctx->d_flat, (uint32_t *)d_streams, cap_words, ctx->d_dirE, ctx->d_err,
                      ctx->d_status, ctx->d_ticket, ctx->ticket_base, ctx->fused_epoch, s); }
        if (exact) {
            ctx->ticket_base += (uint64_t)n * g.nb;  // one ticket per CTA, one CTA per block
            KernelTimer t(ctx, FLIC_K_FINALIZE, s);
            launch_finalize(g, ctx->d_dirE, (uint32_t *)d_streams, cap_words, (unsigned long long *)d_offsets, ctx->d_err, s);
        }
        ctx->launches += exact ? 4 : (fused ? 4 : 5);
    }
    CU(cudaEventRecord(ctx->ev_ws, s));
    ctx->ws_stream = s; ctx->ws_used = true;
    CU(cudaGetLastError());
    return FLIC_OK;
}

// TMA descriptor of a tightly packed pixel batch: 3-D {row elements, rows, images}, boxes of box_elems x 32 rows.
// Returns false when the layout does not qualify (then the kernels load / store directly) or the driver entry
// point is missing.
static bool make_tile_map(const Geo &g, const uint8_t *d_pixels, CUtensorMapDataType dt, uint32_t elem_bytes, uint32_t box_elems,
                          CUtensorMapSwizzle swz, CUtensorMap *tm) {
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
    if (off || !encode || !g.aligned16 || box_elems > 256) return false;
    const cuuint64_t dims[3] = {g.pitch / elem_bytes, g.h, g.n};
    const cuuint64_t strides[2] = {g.pitch, g.img_stride};
    const cuuint32_t box[3] = {box_elems, (cuuint32_t)kBH, 1}, estr[3] = {1, 1, 1};
    return encode(tm, dt, 3, const_cast<uint8_t *>(d_pixels), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int flic_decode_batch_device(flic_ctx *ctx, const uint8_t *d_streams, const uint64_t *d_offsets, uint32_t n,
                                        uint32_t w, uint32_t h, uint32_t c, uint32_t flags, uint8_t *d_pixels,
                                        void *stream) {
    if (!ctx || !d_pixels || !d_streams || !d_offsets || ((uintptr_t)d_streams & 3u)) return FLIC_E_ARG;
    Geo g;
    int rc = make_geo(d_pixels, n, w, h, c, flags, &g);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if (flags & FLIC_FLAG_ONE_STREAM) {
        KernelTimer t(ctx, FLIC_K_DECODE_ONE, (cudaStream_t)stream);
        launch_decode_one((const uint32_t *)d_streams, (const unsigned long long *)d_offsets, g, d_pixels, ctx->d_err + 1,
                          ctx->d_phase ? ctx->d_phase + FLIC_PHASES : nullptr, (cudaStream_t)stream);
    } else {
        KernelTimer t(ctx, FLIC_K_DECODE, (cudaStream_t)stream);
        alignas(64) CUtensorMap tm;
        // TMA tile stores: RGBA in 64-byte rows with the 64 B swizzle, RGB in 96-byte rows (FLIC_NO_TMA_RGB=1: A/B switch)
        static const bool no_rgb = getenv("FLIC_NO_TMA_RGB") != nullptr;
        const bool tma = (g.c == 4 && make_tile_map(g, d_pixels, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 64, CU_TENSOR_MAP_SWIZZLE_64B, &tm)) ||
                         (g.c == 3 && !no_rgb && make_tile_map(g, d_pixels, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 96, CU_TENSOR_MAP_SWIZZLE_NONE, &tm));
        launch_decode((const uint32_t *)d_streams, (const unsigned long long *)d_offsets, g, d_pixels, ctx->d_err + 1,
                      tma ? &tm : nullptr, (cudaStream_t)stream);
    }
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
    int rc = FLIC_OK;
    std::lock_guard<std::mutex> lk(ctx->span_mu);
    for (auto &sp : ctx->spans) {
        float t = 0.f;
        cudaError_t e = cudaEventSynchronize(sp.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, sp.a, sp.b);
        if (e != cudaSuccess) { rc = cuda_fail(ctx, e, "kernel timing events"); }
        else { ms[sp.kernel] += t; counts[sp.kernel] += 1; }
        ctx->free_spans.push_back(sp);
    }
    ctx->spans.clear();
    return rc;
}

static int report_device_errors(flic_ctx *ctx, uint32_t e) {
    if (!e) return FLIC_OK;
    snprintf(ctx->msg, sizeof ctx->msg, "device error bits 0x%x%s%s%s%s%s", e, (e & kErrCapacity) ? " capacity" : "",
             (e & kErrSlot) ? " slot-overrun" : "", (e & kErrFormat) ? " format" : "", (e & kErrRange) ? " payload-exceeds-u32-words" : "",
             (e & kErrLayout) ? " shared-memory-layout" : "");
    if (e & kErrFormat) return FLIC_E_FORMAT;
    if (e & kErrCapacity) return FLIC_E_CAPACITY;
    if (e & kErrRange) return FLIC_E_UNSUPPORTED;
    return FLIC_E_INTERNAL;
}

// One direction's error word (0 encode, 1 decode), read and cleared on `s` through that direction's own pinned word.
static int check_word(flic_ctx *ctx, int word, cudaStream_t s) {
    CU(cudaMemcpyAsync(ctx->h_err + word, ctx->d_err + word, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaMemsetAsync(ctx->d_err + word, 0, sizeof(uint32_t), s));
    CU(cudaStreamSynchronize(s));
    return report_device_errors(ctx, ctx->h_err[word]);
}

extern "C" int flic_get_phase_clocks(flic_ctx *ctx, uint64_t cycles[2 * FLIC_PHASES]) {
    if (!ctx || !cycles) return FLIC_E_ARG;
    if (!ctx->d_phase) return FLIC_E_UNSUPPORTED;  // the context was created without FLIC_PHASE_CLOCKS=1
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(cycles, ctx->d_phase, 2 * FLIC_PHASES * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(cudaMemset(ctx->d_phase, 0, 2 * FLIC_PHASES * sizeof(unsigned long long)));
    return FLIC_OK;
}

extern "C" int flic_check(flic_ctx *ctx, void *stream) {
    if (!ctx) return FLIC_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->h_err + 2, ctx->d_err, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaMemsetAsync(ctx->d_err, 0, 2 * sizeof(uint32_t), s));
    CU(cudaStreamSynchronize(s));
    return report_device_errors(ctx, ctx->h_err[2] | ctx->h_err[3]);
}

// ------------------------------------------------------------ host-buffer API
extern "C" int flic_host_register(void *p, uint64_t bytes) {
    if (!p || !bytes) return FLIC_E_ARG;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return FLIC_OK; }
    return e == cudaSuccess ? FLIC_OK : FLIC_E_CUDA;
}
extern "C" int flic_host_unregister(void *p) {
    if (!p) return FLIC_E_ARG;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) cudaGetLastError();
    return e == cudaSuccess ? FLIC_OK : FLIC_E_CUDA;
}

namespace {
// Pins the caller's two buffers for the duration of one host-API call when they are pageable: cudaMemcpyAsync on
// pageable memory is staged through the driver's bounce buffer and serialises with the host, which turns the
// three-stream pipeline into a sequence.  Buffers that are already pinned (cudaHostAlloc / flic_host_register —
// what a caller on the fast path should do once, up front) are left alone; a failed registration just means the
// slow copies.  Registration is by whole pages: when the two buffers share a page (neighbours on the heap) they
// are registered as ONE range — two overlapping registrations would leave one buffer half pinned, and a copy
// that straddles pinned and pageable memory is an error.
struct AutoPin {
    void *base[2] = {nullptr, nullptr};
    static bool pageable(const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeUnregistered;
    }
    void pin(int slot, uintptr_t lo, uintptr_t hi) {
        if (cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterPortable) == cudaSuccess) base[slot] = (void *)lo;
        else cudaGetLastError();
    }
    AutoPin(const void *p0, uint64_t n0, const void *p1, uint64_t n1) {
        static const bool off = getenv("FLIC_NO_AUTOPIN") != nullptr;
        if (off) return;
        const uintptr_t pg = 4095;
        uintptr_t lo[2] = {(uintptr_t)p0 & ~pg, (uintptr_t)p1 & ~pg};
        uintptr_t hi[2] = {((uintptr_t)p0 + n0 + pg) & ~pg, ((uintptr_t)p1 + n1 + pg) & ~pg};
        // small buffers: registration costs more than it saves
        bool want[2] = {p0 && n0 >= (1u << 20) && pageable(p0), p1 && n1 >= (1u << 20) && pageable(p1)};
        if (want[0] && want[1] && lo[0] < hi[1] && lo[1] < hi[0]) {  // the page ranges overlap: one registration
            pin(0, lo[0] < lo[1] ? lo[0] : lo[1], hi[0] > hi[1] ? hi[0] : hi[1]);
            return;
        }
        for (int i = 0; i < 2; ++i)
            if (want[i]) pin(i, lo[i], hi[i]);
    }
    ~AutoPin() {
        for (int i = 0; i < 2; ++i)
            if (base[i]) cudaHostUnregister(base[i]);
    }
};
}  // namespace

static int ensure_staging(flic_ctx *ctx, Pipe &p, uint64_t pix, uint64_t str, uint64_t noff) {
    if (pix > p.pix_cap) {
        for (int i = 0; i < kDepth; ++i) { cudaFree(p.d_pix[i]); p.d_pix[i] = nullptr; }
        p.pix_cap = 0;
        for (int i = 0; i < kDepth; ++i) CU(cudaMalloc(&p.d_pix[i], pix));
        p.pix_cap = pix;
    }
    if (str > p.str_cap) {
        for (int i = 0; i < kDepth; ++i) { cudaFree(p.d_str[i]); p.d_str[i] = nullptr; }
        p.str_cap = 0;
        for (int i = 0; i < kDepth; ++i) CU(cudaMalloc(&p.d_str[i], str + 16));
        p.str_cap = str;
    }
    if (noff > p.off_cap) {
        for (int i = 0; i < kDepth; ++i) {
            cudaFree(p.d_off[i]); p.d_off[i] = nullptr;
            if (p.h_off[i]) cudaFreeHost(p.h_off[i]);
            p.h_off[i] = nullptr;
        }
        p.off_cap = 0;
        for (int i = 0; i < kDepth; ++i) {
            CU(cudaMalloc(&p.d_off[i], noff * sizeof(unsigned long long)));
            CU(cudaHostAlloc(&p.h_off[i], noff * sizeof(unsigned long long), cudaHostAllocMapped));
            CU(cudaHostGetDevicePointer(&p.m_off[i], p.h_off[i], 0));
        }
        p.off_cap = noff;
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

static void drain(Pipe &p) {  // after a failure: leave no work in flight on the staging or the caller's buffers
    cudaStreamSynchronize(p.s_in); cudaStreamSynchronize(p.s_k); cudaStreamSynchronize(p.s_out);
}

static int encode_batch_impl(flic_ctx *ctx, const uint8_t *h_pixels, uint32_t n, uint32_t w, uint32_t h, uint32_t c,
                             uint32_t flags, uint8_t *h_streams, uint64_t capacity_bytes, uint64_t *h_offsets) {
    Pipe &P = ctx->enc;
    const uint64_t img_bytes = (uint64_t)w * h * c, img_worst = flic_max_stream_bytes(w, h, c);
    const uint32_t m = chunk_images(n, img_bytes);
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_staging(ctx, P, m * img_bytes, m * img_worst, (uint64_t)m + 1);
    if (rc) return rc;
    AutoPin pin(h_pixels, (uint64_t)n * img_bytes, h_streams, capacity_bytes);
    const uint32_t chunks = (n + m - 1) / m;
    uint64_t out_pos = 0;
    h_offsets[0] = 0;
    // Software pipeline over chunks, three host stages per chunk:
    //   A  H2D of the chunk's pixels — issued once the host has seen the kernels of the chunk that last used the
    //      staging buffer finish, up to kDepth chunks ahead of the kernels
    //   B  kernels (their wait for the H2D is on the compute queue, where it stalls nobody else); the stream
    //      offsets land in mapped pinned memory
    //   C  wait for the kernels, place the chunk in the caller's buffer, D2H of exactly that many bytes
    uint32_t issued_in = 0, launched = 0, retired = 0;
    auto stage_a = [&]() -> int {
        const uint32_t k = issued_in++;
        const int b = k % kDepth;
        const uint32_t first = k * m, cnt = (first + m <= n) ? m : n - first;
        CU(cudaMemcpyAsync(P.d_pix[b], h_pixels + (uint64_t)first * img_bytes, cnt * img_bytes, cudaMemcpyHostToDevice, P.s_in));
        CU(cudaEventRecord(P.ev_in[b], P.s_in));
        return FLIC_OK;
    };
    auto stage_b = [&](uint32_t k) -> int {
        const int b = k % kDepth;
        const uint32_t first = k * m, cnt = (first + m <= n) ? m : n - first;
        CU(cudaStreamWaitEvent(P.s_k, P.ev_in[b], 0));
        if (k >= (uint32_t)kDepth) CU(cudaStreamWaitEvent(P.s_k, P.ev_out[b], 0));  // D2H of chunk k-kDepth has drained d_str[b]
        int r = flic_encode_batch_device(ctx, P.d_pix[b], cnt, w, h, c, flags, P.d_str[b], cnt * img_worst, (uint64_t *)P.m_off[b], P.s_k);
        if (r) return r;
        CU(cudaEventRecord(P.ev_k[b], P.s_k));
        return FLIC_OK;
    };
    auto stage_c = [&](uint32_t k) -> int {
        const int b = k % kDepth;
        const uint32_t first = k * m, cnt = (first + m <= n) ? m : n - first;
        CU(wait_chunk(P.ev_k[b], issued_in - k <= 1));
        const uint64_t total = P.h_off[b][cnt];
        if (total > cnt * img_worst) return FLIC_E_INTERNAL;  // kernels flagged a capacity overrun
        if (out_pos + total > capacity_bytes) return FLIC_E_CAPACITY;
        for (uint32_t i = 1; i <= cnt; ++i) h_offsets[first + i] = out_pos + P.h_off[b][i];
        CU(cudaMemcpyAsync(h_streams + out_pos, P.d_str[b], total, cudaMemcpyDeviceToHost, P.s_out));  // its kernels are done: runnable now
        CU(cudaEventRecord(P.ev_out[b], P.s_out));
        out_pos += total;
        return FLIC_OK;
    };
    Pipe &other = ctx->dec;
    P.active.store(1);
    while (rc == FLIC_OK && retired < chunks) {
        const uint32_t lim = other.active.load() ? (uint32_t)pipe_depth_both() : (uint32_t)pipe_depth();
        while (rc == FLIC_OK && issued_in < chunks && issued_in - retired < lim) rc = stage_a();  // the buffers of retired chunks are free
        while (rc == FLIC_OK && launched < issued_in) rc = stage_b(launched++);
        if (rc == FLIC_OK) rc = stage_c(retired++);
    }
    P.active.store(0);
    drain(P);  // success or not: nothing may touch the caller's buffers after the call returns
    const int chk = check_word(ctx, 0, P.s_k);
    return rc ? rc : chk;
}

static int encode_args_ok(flic_ctx *ctx, const uint8_t *h_pixels, uint32_t n, uint32_t w, uint32_t h, uint32_t c,
                          uint32_t flags, uint8_t *h_streams, uint64_t *h_offsets) {
    if (!ctx || !h_pixels || !h_streams || !h_offsets) return FLIC_E_ARG;
    if (n == 0 || w == 0 || h == 0 || c < 1 || c > 4 || !flags_ok(flags)) return FLIC_E_ARG;
    if (cdiv(w, kBW) * cdiv(h, kBH) > 0xFFFFFFFFull) return FLIC_E_ARG;
    return FLIC_OK;
}

extern "C" int flic_encode_batch(flic_ctx *ctx, const uint8_t *h_pixels, uint32_t n, uint32_t w, uint32_t h, uint32_t c,
                                 uint32_t flags, uint8_t *h_streams, uint64_t capacity_bytes, uint64_t *h_offsets) {
    int rc = encode_args_ok(ctx, h_pixels, n, w, h, c, flags, h_streams, h_offsets);
    if (rc) return rc;
    if (ctx->enc.busy) return FLIC_E_BUSY;
    return encode_batch_impl(ctx, h_pixels, n, w, h, c, flags, h_streams, capacity_bytes, h_offsets);
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
    if (!flags_ok(info->flags)) return FLIC_E_FORMAT;
    if (info->block_w == 0 || info->block_h == 0 || (info->block_h & 1)) return FLIC_E_FORMAT;
    if (cdiv(info->width, info->block_w) * cdiv(info->height, info->block_h) != info->n_blocks) return FLIC_E_FORMAT;
    if (4ull * (kHdrWords + (uint64_t)info->n_blocks + 1 + info->payload_words) > size) return FLIC_E_FORMAT;
    return FLIC_OK;
}

// Decodes the streams [lo, hi) of a batch, all of one geometry `g0`, through the decode pipe.
static int decode_run(flic_ctx *ctx, const uint8_t *h_streams, const uint64_t *h_offsets, uint32_t lo, uint32_t hi,
                      const flic_image_info &g0, uint8_t *h_pixels) {
    Pipe &P = ctx->dec;
    const uint32_t n = hi - lo;
    const uint64_t img_bytes = (uint64_t)g0.width * g0.height * g0.channels;
    const uint32_t m = chunk_images(n, img_bytes);
    const uint32_t chunks = (n + m - 1) / m;
    uint64_t max_str = 0;
    for (uint32_t k = 0; k < chunks; ++k) {
        const uint32_t f0 = lo + k * m, f1 = (f0 + m <= hi) ? f0 + m : hi;
        const uint64_t sz = h_offsets[f1] - h_offsets[f0];
        if (sz > max_str) max_str = sz;
    }
    int rc = ensure_staging(ctx, P, m * img_bytes, max_str, (uint64_t)m + 1);
    if (rc) return rc;
    // Same rule as the encoder's pipeline: the H2D of a chunk is issued when the host has seen the kernel that last
    // read the staging buffer finish; the D2H of a chunk is issued when the host has seen its kernel finish.
    uint32_t issued = 0, retired = 0;
    auto issue = [&](uint32_t k) -> int {
        const int b = k % kDepth;
        const uint32_t f0 = lo + k * m, cnt = (f0 + m <= hi) ? m : hi - f0;
        const uint64_t base = h_offsets[f0], sz = h_offsets[f0 + cnt] - base;
        for (uint32_t i = 0; i <= cnt; ++i) P.h_off[b][i] = h_offsets[f0 + i] - base;
        CU(cudaMemcpyAsync(P.d_str[b], h_streams + base, sz, cudaMemcpyHostToDevice, P.s_in));
        CU(cudaMemcpyAsync(P.d_off[b], P.h_off[b], ((uint64_t)cnt + 1) * 8, cudaMemcpyHostToDevice, P.s_in));
        CU(cudaEventRecord(P.ev_in[b], P.s_in));
        CU(cudaStreamWaitEvent(P.s_k, P.ev_in[b], 0));
        if (k >= (uint32_t)kDepth) CU(cudaStreamWaitEvent(P.s_k, P.ev_out[b], 0));  // D2H of chunk k-kDepth has drained d_pix[b]
        int r = flic_decode_batch_device(ctx, P.d_str[b], (const uint64_t *)P.d_off[b], cnt, g0.width, g0.height, g0.channels,
                                         g0.flags, P.d_pix[b], P.s_k);
        if (r) return r;
        CU(cudaEventRecord(P.ev_k[b], P.s_k));
        return FLIC_OK;
    };
    auto retire = [&](uint32_t k) -> int {
        const int b = k % kDepth;
        const uint32_t f0 = lo + k * m, cnt = (f0 + m <= hi) ? m : hi - f0;
        CU(wait_chunk(P.ev_k[b], issued - k <= 1));  // the chunk's pixels exist: its D2H is runnable, its stream buffer is free
        CU(cudaMemcpyAsync(h_pixels + (uint64_t)(f0 - lo) * img_bytes, P.d_pix[b], cnt * img_bytes, cudaMemcpyDeviceToHost, P.s_out));
        CU(cudaEventRecord(P.ev_out[b], P.s_out));
        return FLIC_OK;
    };
    Pipe &other = ctx->enc;
    P.active.store(1);
    while (rc == FLIC_OK && retired < chunks) {
        const uint32_t lim = other.active.load() ? (uint32_t)pipe_depth_both() : (uint32_t)pipe_depth();
        while (rc == FLIC_OK && issued < chunks && issued - retired < lim) rc = issue(issued++);
        if (rc == FLIC_OK) rc = retire(retired++);
    }
    P.active.store(0);
    drain(P);
    return rc;
}

static bool same_geometry(const flic_image_info &a, const flic_image_info &b) {
    return a.width == b.width && a.height == b.height && a.channels == b.channels &&
           ((a.flags ^ b.flags) & ~FLIC_FLAG_EXACT) == 0;  // EXACT does not change how a stream decodes
}

static int decode_batch_impl(flic_ctx *ctx, const uint8_t *h_streams, const uint64_t *h_offsets, uint32_t n,
                             uint8_t *h_pixels, uint64_t pixels_capacity) {
    // A batch may mix geometries: it is decoded as runs of consecutive streams of one geometry, one launch
    // sequence per run, pixels tightly packed in stream order.
    std::vector<flic_image_info> infos(n);
    uint64_t need = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (h_offsets[i + 1] < h_offsets[i] || (h_offsets[i] & 3u)) return FLIC_E_FORMAT;
        int rc = flic_peek(h_streams + h_offsets[i], h_offsets[i + 1] - h_offsets[i], &infos[i]);
        if (rc) return rc;
        if (infos[i].block_w != (uint32_t)kBW || infos[i].block_h != (uint32_t)kBH) return FLIC_E_UNSUPPORTED;
        need += (uint64_t)infos[i].width * infos[i].height * infos[i].channels;
    }
    if (need > pixels_capacity) return FLIC_E_CAPACITY;
    CU(cudaSetDevice(ctx->device));
    AutoPin pin(h_streams + h_offsets[0], h_offsets[n] - h_offsets[0], h_pixels, need);
    int rc = FLIC_OK;
    uint64_t pix_pos = 0;
    for (uint32_t lo = 0; lo < n && rc == FLIC_OK;) {
        uint32_t hi = lo + 1;
        while (hi < n && same_geometry(infos[hi], infos[lo])) ++hi;
        rc = decode_run(ctx, h_streams, h_offsets, lo, hi, infos[lo], h_pixels + pix_pos);
        pix_pos += (uint64_t)(hi - lo) * infos[lo].width * infos[lo].height * infos[lo].channels;
        lo = hi;
    }
    const int chk = check_word(ctx, 1, ctx->dec.s_k);
    return rc ? rc : chk;
}

extern "C" int flic_decode_batch(flic_ctx *ctx, const uint8_t *h_streams, const uint64_t *h_offsets, uint32_t n,
                                 uint8_t *h_pixels, uint64_t pixels_capacity) {
    if (!ctx || !h_streams || !h_offsets || !h_pixels || n == 0) return FLIC_E_ARG;
    if (ctx->dec.busy) return FLIC_E_BUSY;
    return decode_batch_impl(ctx, h_streams, h_offsets, n, h_pixels, pixels_capacity);
}

// ---- submit / wait: the same pipelines on a worker thread, so that one call's D2H overlaps another's H2D ----
extern "C" int flic_encode_submit(flic_ctx *ctx, const uint8_t *h_pixels, uint32_t n, uint32_t w, uint32_t h, uint32_t c,
                                  uint32_t flags, uint8_t *h_streams, uint64_t capacity_bytes, uint64_t *h_offsets) {
    int rc = encode_args_ok(ctx, h_pixels, n, w, h, c, flags, h_streams, h_offsets);
    if (rc) return rc;
    Pipe &P = ctx->enc;
    if (P.busy) return FLIC_E_BUSY;
    P.busy = true;
    P.worker = std::thread([=, &P] { P.result = encode_batch_impl(ctx, h_pixels, n, w, h, c, flags, h_streams, capacity_bytes, h_offsets); });
    return FLIC_OK;
}

extern "C" int flic_decode_submit(flic_ctx *ctx, const uint8_t *h_streams, const uint64_t *h_offsets, uint32_t n,
                                  uint8_t *h_pixels, uint64_t pixels_capacity) {
    if (!ctx || !h_streams || !h_offsets || !h_pixels || n == 0) return FLIC_E_ARG;
    Pipe &P = ctx->dec;
    if (P.busy) return FLIC_E_BUSY;
    P.busy = true;
    P.worker = std::thread([=, &P] { P.result = decode_batch_impl(ctx, h_streams, h_offsets, n, h_pixels, pixels_capacity); });
    return FLIC_OK;
}

extern "C" int flic_wait(flic_ctx *ctx, int op) {
    if (!ctx || (op != FLIC_OP_ENCODE && op != FLIC_OP_DECODE)) return FLIC_E_ARG;
    Pipe &P = op == FLIC_OP_ENCODE ? ctx->enc : ctx->dec;
    if (!P.busy) return FLIC_E_ARG;
    P.worker.join();
    P.busy = false;
    return P.result;
}

// ------------------------------------------------------------------- splice
extern "C" int flic_splice_plan(const uint32_t *part_blocks, const uint32_t *part_payload_words, uint32_t k,
                                uint64_t *dir_byte_off, uint64_t *payload_byte_off, uint64_t *total_bytes) {
    if (!part_blocks || !part_payload_words || !dir_byte_off || !payload_byte_off || !total_bytes || k == 0) return FLIC_E_ARG;
    uint64_t nb = 0, pw = 0;
    for (uint32_t i = 0; i < k; ++i) { nb += part_blocks[i]; pw += part_payload_words[i]; }
    if (nb >= (1ull << 32) || pw >= (1ull << 32)) return FLIC_E_ARG;
    uint64_t b = 0, p = 0;
    for (uint32_t i = 0; i < k; ++i) {
        dir_byte_off[i] = 4ull * (kHdrWords + b);
        payload_byte_off[i] = 4ull * (kHdrWords + nb + 1 + p);
        b += part_blocks[i]; p += part_payload_words[i];
    }
    *total_bytes = 4ull * (kHdrWords + nb + 1 + pw);
    return FLIC_OK;
}

extern "C" int flic_splice_finish_device(flic_ctx *ctx, uint8_t *d_out, const uint32_t *part_blocks,
                                         const uint32_t *part_payload_words, uint32_t k, uint32_t w, uint32_t h_total,
                                         uint32_t c, uint32_t flags, void *stream) {
    if (!ctx || !d_out || ((uintptr_t)d_out & 3u) || !part_blocks || !part_payload_words || k == 0 || k > FLIC_MAX_PARTS) return FLIC_E_ARG;
    if (w == 0 || h_total == 0 || c < 1 || c > 4 || !flags_ok(flags)) return FLIC_E_ARG;
    SpliceParts sp;
    uint64_t nb = 0, pw = 0;
    for (uint32_t i = 0; i < k; ++i) {
        sp.first_block[i] = (uint32_t)nb; sp.base_words[i] = (uint32_t)pw;
        nb += part_blocks[i]; pw += part_payload_words[i];
    }
    if (nb >= (1ull << 32) || pw >= (1ull << 32) || nb != cdiv(w, kBW) * cdiv(h_total, kBH)) return FLIC_E_ARG;
    sp.first_block[k] = (uint32_t)nb; sp.base_words[k] = (uint32_t)pw;
    sp.k = k;
    CU(cudaSetDevice(ctx->device));
    launch_splice_finish((uint32_t *)d_out, sp, w, h_total, c, flags, (cudaStream_t)stream);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

// ---- block-row split with peer memory: the staged encoder in two halves, with the one collective — the
// all-gather of the parts' payload sizes — between them, and no host synchronisation anywhere.
extern "C" int flic_encode_plan_device(flic_ctx *ctx, const uint8_t *d_pixels, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                                       uint64_t *d_payload_words, void *stream) {
    if (!ctx || !d_pixels || !d_payload_words) return FLIC_E_ARG;
    if (flags & (FLIC_FLAG_ONE_STREAM | FLIC_FLAG_EXACT)) return FLIC_E_UNSUPPORTED;  // positions must follow from the histograms
    Geo g;
    int rc = make_geo(d_pixels, 1, w, h, c, flags, &g);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    rc = ensure_workspace(ctx, g.nb, true);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (ctx->ws_used && ctx->ws_stream != s) CU(cudaStreamWaitEvent(s, ctx->ev_ws, 0));
    { KernelTimer t(ctx, FLIC_K_HISTOGRAMS, s);
      alignas(64) CUtensorMap tm;
      const bool tma = make_load_map(g, d_pixels, &tm);
      launch_histograms(d_pixels, g, ctx->d_hist, ctx->d_resid, ctx->d_flat, tma ? &tm : nullptr, s); }
    { KernelTimer t(ctx, FLIC_K_TABLES, s); launch_tables(ctx->d_hist, g.nb, ctx->d_table, ctx->d_bits, s); }
    { KernelTimer t(ctx, FLIC_K_SLOTS, s);
      launch_slots(g, ctx->d_bits, ctx->d_dirE, ctx->d_slot_status, ++ctx->slot_epoch, ~0ull, ctx->d_err, nullptr, nullptr,
                   ctx->slots_max_grid, s); }
    launch_part_words(ctx->d_dirE, g.nb, (unsigned long long *)d_payload_words, s);
    ctx->launches += 4;
    ctx->plan_geo = g; ctx->plan_valid = true;
    CU(cudaEventRecord(ctx->ev_ws, s));
    ctx->ws_stream = s; ctx->ws_used = true;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_encode_emit_device(flic_ctx *ctx, uint8_t *d_stream, uint64_t capacity_bytes, uint32_t total_blocks,
                                       uint32_t first_block, const uint64_t *d_base_words, void *stream) {
    if (!ctx || !d_stream || !d_base_words || ((uintptr_t)d_stream & 3u)) return FLIC_E_ARG;
    if (!ctx->plan_valid) return FLIC_E_ARG;
    const Geo g = ctx->plan_geo;
    if ((uint64_t)first_block + g.nb > total_blocks) return FLIC_E_ARG;
    const uint64_t hdr_words = (uint64_t)kHdrWords + total_blocks + 1;
    if (capacity_bytes / 4 < hdr_words || hdr_words > 0xFFFFFFFFull) return FLIC_E_CAPACITY;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (ctx->ws_used && ctx->ws_stream != s) CU(cudaStreamWaitEvent(s, ctx->ev_ws, 0));
    launch_part_directory(ctx->d_dirE, g.nb, (const unsigned long long *)d_base_words, (uint32_t *)d_stream + kHdrWords + first_block, s);
    { KernelTimer t(ctx, FLIC_K_PACK, s);
      launch_pack(ctx->d_resid, g, ctx->d_table, ctx->d_flat, (uint32_t *)d_stream, capacity_bytes / 4, ctx->d_dirE, ctx->d_err,
                  ctx->d_status, ctx->d_ticket, ctx->ticket_base, ctx->fused_epoch, s, (const unsigned long long *)d_base_words,
                  (uint32_t)hdr_words); }
    ctx->launches += 2;
    ctx->plan_valid = false;
    CU(cudaEventRecord(ctx->ev_ws, s));
    ctx->ws_stream = s; ctx->ws_used = true;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_splice_header_device(flic_ctx *ctx, uint8_t *d_stream, uint64_t capacity_bytes, uint32_t w, uint32_t h_total,
                                         uint32_t c, uint32_t flags, const uint64_t *d_total_words, void *stream) {
    if (!ctx || !d_stream || !d_total_words || ((uintptr_t)d_stream & 3u) || w == 0 || h_total == 0 || c < 1 || c > 4 || !flags_ok(flags))
        return FLIC_E_ARG;
    const uint64_t nb = cdiv(w, kBW) * cdiv(h_total, kBH);
    if (nb >= (1ull << 32)) return FLIC_E_ARG;
    CU(cudaSetDevice(ctx->device));
    launch_splice_header((uint32_t *)d_stream, w, h_total, c, flags, (uint32_t)nb, (const unsigned long long *)d_total_words,
                         capacity_bytes / 4, ctx->d_err, (cudaStream_t)stream);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_pull_part_device(flic_ctx *ctx, const uint8_t *d_stream, uint64_t stream_bytes, uint32_t total_blocks,
                                     uint32_t first_block, uint32_t part_blocks, uint8_t *d_part, uint64_t capacity_bytes,
                                     uint64_t *d_part_bytes, void *stream) {
    if (!ctx || !d_stream || !d_part || !d_part_bytes || (((uintptr_t)d_stream | (uintptr_t)d_part) & 3u)) return FLIC_E_ARG;
    if ((uint64_t)first_block + part_blocks > total_blocks) return FLIC_E_ARG;
    CU(cudaSetDevice(ctx->device));
    launch_pull_part((const uint32_t *)d_stream, stream_bytes / 4, total_blocks, first_block, part_blocks, (uint32_t *)d_part, capacity_bytes / 4,
                     (unsigned long long *)d_part_bytes, ctx->d_err + 1, (cudaStream_t)stream);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_split_finish_device(flic_ctx *ctx, uint8_t *d_part, uint32_t w, uint32_t h_part, uint32_t c,
                                        uint32_t flags, void *stream) {
    if (!ctx || !d_part || ((uintptr_t)d_part & 3u) || w == 0 || h_part == 0 || c < 1 || c > 4 || !flags_ok(flags)) return FLIC_E_ARG;
    const uint64_t nb = cdiv(w, kBW) * cdiv(h_part, kBH);
    if (nb >= (1ull << 32)) return FLIC_E_ARG;
    CU(cudaSetDevice(ctx->device));
    launch_split_finish((uint32_t *)d_part, (uint32_t)nb, w, h_part, c, flags, (cudaStream_t)stream);
    ctx->launches += 1;
    CU(cudaGetLastError());
    return FLIC_OK;
}

extern "C" int flic_splice_block_rows_device(flic_ctx *ctx, const uint8_t *const *d_parts, const uint64_t *part_sizes,
                                             uint32_t k, uint8_t *d_out, uint64_t out_capacity, uint64_t *out_size, void *stream) {
    if (!ctx || !d_parts || !part_sizes || !d_out || !out_size || k == 0 || k > FLIC_MAX_PARTS) return FLIC_E_ARG;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t hdr[FLIC_MAX_PARTS][FLIC_HEADER_BYTES];
    for (uint32_t i = 0; i < k; ++i) {
        if (!d_parts[i] || part_sizes[i] < FLIC_HEADER_BYTES || ((uintptr_t)d_parts[i] & 3u)) return FLIC_E_ARG;
        CU(cudaMemcpyAsync(hdr[i], d_parts[i], FLIC_HEADER_BYTES, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    flic_image_info first, info;
    uint32_t nbs[FLIC_MAX_PARTS], pws[FLIC_MAX_PARTS];
    uint64_t height = 0;
    for (uint32_t i = 0; i < k; ++i) {
        // the header alone is on the host: check it against the part's size without touching the payload
        memset(&info, 0, sizeof info);
        uint8_t tmp[FLIC_HEADER_BYTES];
        memcpy(tmp, hdr[i], sizeof tmp);
        uint32_t wd[8];
        memcpy(wd, tmp, sizeof wd);
        if (4ull * (kHdrWords + (uint64_t)wd[5] + 1 + wd[6]) > part_sizes[i]) return FLIC_E_FORMAT;
        int rc = flic_peek(tmp, part_sizes[i], &info);
        if (rc) return rc;
        if (i == 0) first = info;
        else if (info.width != first.width || info.channels != first.channels || info.flags != first.flags ||
                 info.block_w != first.block_w || info.block_h != first.block_h)
            return FLIC_E_ARG;
        if (i + 1 < k && info.height % info.block_h) return FLIC_E_ARG;  // only the last part may be ragged
        nbs[i] = info.n_blocks; pws[i] = info.payload_words; height += info.height;
    }
    if (height >= (1ull << 32)) return FLIC_E_ARG;
    uint64_t doff[FLIC_MAX_PARTS], poff[FLIC_MAX_PARTS], total = 0;
    int rc = flic_splice_plan(nbs, pws, k, doff, poff, &total);
    if (rc) return rc;
    if (total > out_capacity) return FLIC_E_CAPACITY;
    for (uint32_t i = 0; i < k; ++i) {
        const uint8_t *pdir = d_parts[i] + 4 * kHdrWords;
        CU(cudaMemcpyAsync(d_out + doff[i], pdir, 4ull * nbs[i], cudaMemcpyDeviceToDevice, s));
        CU(cudaMemcpyAsync(d_out + poff[i], pdir + 4ull * (nbs[i] + 1), 4ull * pws[i], cudaMemcpyDeviceToDevice, s));
    }
    rc = flic_splice_finish_device(ctx, d_out, nbs, pws, k, first.width, (uint32_t)height, first.channels, first.flags, stream);
    if (rc) return rc;
    *out_size = total;
    return FLIC_OK;
}

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
