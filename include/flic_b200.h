/*
 * flic_b200.h — C ABI of the B200 block-codec engine (libflicb200.so).
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types in any signature
 * (streams travel as void* holding a cudaStream_t).  Every entry point returns
 * 0 (FLIC_OK) or a negative FLIC_E_* code; the library never falls back to a
 * CPU path — with no usable sm_100 device, flic_create() fails.
 *
 * WHAT THIS REPLACES IN THE REFERENCE: **not stated, by necessity.**  The spec
 * (BASELINE.json north_star) places the drop-in boundary at the reference's
 * Rust encode/decode entry points in src/image.rs / src/main.rs.  Those files
 * are behind the licensing gate (LICENSING.md) and were not read, so this
 * header cannot cite the signatures it would stand in for.  The boundary below
 * is therefore *defined here*, in the shape SURVEY.md §8(b) prescribes (batch
 * encode/decode, caller-owned buffers, size-query call, integer status codes,
 * one CUDA stream per call, thread-compatible not thread-safe), and
 * INTEGRATION.md shows the generic Rust `extern "C"` block that would bind it.
 * The bitstream it produces is the provisional FLP0 format, NOT
 * the reference's format.
 */
#ifndef FLIC_B200_H
#define FLIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLIC_OK 0
#define FLIC_E_ARG (-1)         /* null pointer, zero dimension, channels not in 1..4, bad flags */
#define FLIC_E_CAPACITY (-2)    /* caller-owned output buffer too small */
#define FLIC_E_FORMAT (-3)      /* stream header / directory inconsistent */
#define FLIC_E_CUDA (-4)        /* a CUDA call failed; see flic_last_error() */
#define FLIC_E_NO_DEVICE (-5)   /* no sm_100 device: there is no CPU fallback */
#define FLIC_E_UNSUPPORTED (-6) /* valid FLP0 feature this build has no kernel for */
#define FLIC_E_INTERNAL (-7)    /* device-side consistency check tripped */
#define FLIC_E_BUSY (-8)        /* a submitted operation of the same kind has not been waited for */

#define FLIC_BLOCK_W 128        /* pixels per block row   (one CTA / one decode warp per block) */
#define FLIC_BLOCK_H 32         /* rows per block         (one decode lane per row sub-stream)  */
#define FLIC_MAX_CODE_LEN 10
#define FLIC_HEADER_BYTES 32

#define FLIC_PRED_LEFT 1u       /* flags bits 0-3: predictor id */
#define FLIC_FLAG_SUBGREEN 0x10u
/* Two optional stream layouts (round 2; both are modes of FLP0 v3 recorded in the header's flags byte,
 * both exist to measure the engine on the PESSIMISTIC assumptions about a format it did not design):
 *  ONE_STREAM  a block is ONE contiguous MSB-first bit stream (all rows back to back, padded to a word
 *              once), with no per-row sub-streams and no row word counts: nothing in the stream tells a
 *              decoder where a row starts, so parallel decode has to self-synchronise (k_decode_one).
 *  EXACT       per-row sub-streams as usual, but a block occupies exactly the words it uses (slack 0):
 *              its size is NOT known before it is packed, so the encoder places blocks with a
 *              decoupled look-back over the packed sizes instead of precomputed slots. */
#define FLIC_FLAG_ONE_STREAM 0x20u
#define FLIC_FLAG_EXACT 0x40u
#define FLIC_FLAGS_ALL 0x7Fu

typedef struct flic_ctx flic_ctx;

typedef struct flic_image_info {
    uint32_t width, height, channels, flags;
    uint32_t block_w, block_h, n_blocks, payload_words;
} flic_image_info;

/* ---- lifetime ---------------------------------------------------------- */
int flic_create(int device, flic_ctx **out);
void flic_destroy(flic_ctx *ctx);
const char *flic_strerror(int code);
const char *flic_last_error(const flic_ctx *ctx); /* detail of the last FLIC_E_CUDA / _INTERNAL */
int flic_version(void);

/* Options.  FLIC_OPT_ENCODER selects the encode path; both produce identical bytes, in every layout.
 *   FUSED   one persistent kernel: pixels are read once, residuals never leave the SM, blocks are placed by a
 *           decoupled look-back (2 launches per call, DRAM traffic = the algorithmic bytes);
 *   STAGED  the five-kernel pipeline with a residual plane in HBM (2.4x the algorithmic DRAM bytes, but every
 *           stage runs at full occupancy and the serial Huffman merges of eight blocks share a warp — on B200
 *           this path is bound by issue slots, not by HBM, and is the faster one for large batches);
 *   AUTO    (default) FUSED for small jobs, where launch count and latency matter, STAGED for large ones.
 * With FLIC_FLAG_ONE_STREAM the staged pipeline's pack kernel concatenates the rows bit-exactly (the block sizes are
 * exact and still follow from the histograms); with FLIC_FLAG_EXACT it takes blocks in ticket order and places them with
 * the same decoupled look-back over the packed sizes as the fused kernel.  The environment variable
 * FLIC_ENCODER=fused|staged|auto sets the default at flic_create(). */
#define FLIC_OPT_ENCODER 1
#define FLIC_ENCODER_FUSED 0
#define FLIC_ENCODER_STAGED 1
#define FLIC_ENCODER_AUTO 2
int flic_set_option(flic_ctx *ctx, int option, int value);

/* ---- size queries ------------------------------------------------------ */
uint64_t flic_blocks_per_image(uint32_t w, uint32_t h);
/* Worst-case bytes of ONE encoded w x h x c image (header + directory + payload). */
uint64_t flic_max_stream_bytes(uint32_t w, uint32_t h, uint32_t c);

/* ---- device-resident batch API (inputs/outputs already in HBM) --------- */
/* Encodes n images of identical geometry, tightly packed at d_pixels
 * (n*h*w*c bytes), into n FLP0 streams laid back to back at d_streams.
 * d_offsets receives n+1 byte offsets (u64, device memory); stream i is
 * [d_offsets[i], d_offsets[i+1]).  capacity_bytes >= n*flic_max_stream_bytes().
 * Asynchronous on `stream`. */
int flic_encode_batch_device(flic_ctx *ctx, const uint8_t *d_pixels, uint32_t n, uint32_t w,
                             uint32_t h, uint32_t c, uint32_t flags, uint8_t *d_streams,
                             uint64_t capacity_bytes, uint64_t *d_offsets, void *stream);

/* Decodes n streams of identical geometry (as produced above) into tightly
 * packed pixels.  Geometry is passed by the caller (the host API reads it
 * from the headers).  Asynchronous on `stream`; device-side format violations
 * surface at the next flic_check(). */
int flic_decode_batch_device(flic_ctx *ctx, const uint8_t *d_streams, const uint64_t *d_offsets,
                             uint32_t n, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                             uint8_t *d_pixels, void *stream);

/* Synchronises `stream` and reports device-side error flags raised by the
 * kernels launched through ctx since the last check (capacity overrun,
 * corrupt directory, slot overrun, an image payload beyond 2^32 words).
 *
 * A context owns ONE encode workspace.  Encodes issued on different streams are
 * ordered by an event (the later one waits for the earlier), so they are safe
 * but do not overlap; use one context per concurrent encode. */
int flic_check(flic_ctx *ctx, void *stream);

/* ---- host-buffer batch API (H2D + kernels + D2H inside the call) -------
 * The batch streams through double-buffered device staging on three CUDA streams (H2D, kernels, D2H).
 * HOST MEMORY: the copies only overlap when the caller's buffers are page-locked.  Allocate them with
 * cudaHostAlloc, or pin them ONCE with flic_host_register() (a cudaHostRegister wrapper for callers that do
 * not link the CUDA runtime).  Pageable buffers of 1 MiB or more are pinned for the duration of each call
 * (cudaHostRegister + cudaHostUnregister: correct, but it costs about 0.2 ms per MiB per call; set
 * FLIC_NO_AUTOPIN=1 to skip it and take the driver's staged copies instead).
 * flic_decode_batch accepts batches of mixed geometry: consecutive streams of one geometry are decoded
 * together, pixels come out tightly packed in stream order. */
int flic_host_register(void *p, uint64_t bytes);
int flic_host_unregister(void *p);
int flic_encode_batch(flic_ctx *ctx, const uint8_t *h_pixels, uint32_t n, uint32_t w, uint32_t h,
                      uint32_t c, uint32_t flags, uint8_t *h_streams, uint64_t capacity_bytes,
                      uint64_t *h_offsets /* n+1 */);
int flic_decode_batch(flic_ctx *ctx, const uint8_t *h_streams, const uint64_t *h_offsets,
                      uint32_t n, uint8_t *h_pixels, uint64_t pixels_capacity);

/* Asynchronous forms: the same pipelines run on a worker thread of the context and the call returns at
 * once; flic_wait(ctx, FLIC_OP_ENCODE / FLIC_OP_DECODE) joins it and returns its status.  Encode and decode
 * own separate staging and streams, so ONE encode and ONE decode may be in flight together (FLIC_E_BUSY
 * for a second of the same kind): an encode's stream download then overlaps a decode's stream upload and
 * both directions of the PCIe link are busy.  The caller's buffers must stay valid, and must not be the
 * other operation's output, until the wait returns.  Device-API calls on the same context must not be
 * issued while an operation is in flight. */
#define FLIC_OP_ENCODE 0
#define FLIC_OP_DECODE 1
int flic_encode_submit(flic_ctx *ctx, const uint8_t *h_pixels, uint32_t n, uint32_t w, uint32_t h,
                       uint32_t c, uint32_t flags, uint8_t *h_streams, uint64_t capacity_bytes,
                       uint64_t *h_offsets /* n+1 */);
int flic_decode_submit(flic_ctx *ctx, const uint8_t *h_streams, const uint64_t *h_offsets,
                       uint32_t n, uint8_t *h_pixels, uint64_t pixels_capacity);
int flic_wait(flic_ctx *ctx, int op);

/* ---- host-only helpers -------------------------------------------------- */
int flic_peek(const uint8_t *stream, uint64_t size, flic_image_info *info);

/* Splices k streams that each hold a run of whole block rows of one image
 * (same width/channels/flags, every part but the last a multiple of
 * FLIC_BLOCK_H rows) into the stream of the full image.  This is the
 * "split one oversized image by block rows" path: blocks never predict across
 * block edges, so the parts' payloads concatenate and only the directory is
 * rebased.  Returns bytes written via *out_size. */
int flic_splice_block_rows(const uint8_t *const *parts, const uint64_t *part_sizes, uint32_t k,
                           uint8_t *out, uint64_t out_capacity, uint64_t *out_size);

/* Device-side splice, for parts that live in HBM (one GPU, or gathered over NVLink).
 * flic_splice_block_rows_device: k part streams in device memory -> one stream at d_out (D2D copies of the
 * directories and payloads + one kernel for the header and the directory rebase); reads the k 32-byte part
 * headers back to the host first.  *out_size = bytes of the spliced stream.
 *
 * The multi-GPU form avoids the staging copy: after the all-gather of every part's (n_blocks,
 * payload_words), flic_splice_plan() says where part i's directory entries (its first n_blocks[i] u32) and
 * its payload land in the spliced stream; each rank sends them straight there (NCCL send/recv or P2P);
 * flic_splice_finish_device() then writes the header, adds each part's payload base to its directory
 * segment and appends the final entry.  flic_split_finish_device() is the inverse for decode: a buffer
 * that received entries [b0, b0+nb] of a stream's directory at byte 32 and the payload words they span right
 * behind them becomes the stand-alone stream of those block rows (header written, directory rebased). */
#define FLIC_MAX_PARTS 64
int flic_splice_block_rows_device(flic_ctx *ctx, const uint8_t *const *d_parts, const uint64_t *part_sizes,
                                  uint32_t k, uint8_t *d_out, uint64_t out_capacity, uint64_t *out_size,
                                  void *stream);
int flic_splice_plan(const uint32_t *part_blocks, const uint32_t *part_payload_words, uint32_t k,
                     uint64_t *dir_byte_off /* k */, uint64_t *payload_byte_off /* k */,
                     uint64_t *total_bytes);
int flic_splice_finish_device(flic_ctx *ctx, uint8_t *d_out, const uint32_t *part_blocks,
                              const uint32_t *part_payload_words, uint32_t k, uint32_t w,
                              uint32_t h_total, uint32_t c, uint32_t flags, void *stream);
int flic_split_finish_device(flic_ctx *ctx, uint8_t *d_part, uint32_t w, uint32_t h_part, uint32_t c,
                             uint32_t flags, void *stream);

/* ---- the same split with PEER MEMORY: pack straight into another GPU's buffer over NVLink ----
 * One process per GPU; d_stream is the spliced stream's buffer, which may live on another GPU and be mapped into
 * this process (CUDA IPC / torch symmetric memory).  No host synchronisation anywhere; the only collective is the
 * all-gather of the parts' payload sizes between the two halves of the encoder:
 *   flic_encode_plan_device   histograms, code tables and slot positions of this GPU's block rows (default layout
 *                             only: positions must follow from the histograms); *d_payload_words (device) = the
 *                             part's payload size in words;
 *   (caller)                  all-gather the sizes, exclusive prefix sum -> *d_base_words (device) per GPU;
 *   flic_encode_emit_device   (FLIC_E_ARG unless a plan is pending: any other encode call on the context discards it)
 *                             packs the planned part: block payloads at word 8 + total_blocks + 1 + *d_base_words +
 *                             (offset within the part) of d_stream, directory entries at word 8 + first_block + i;
 *   flic_splice_header_device (one GPU, after all parts are in: e.g. behind a tiny all-reduce) the 8 header words and
 *                             the final directory entry, from the device-side total;
 *   flic_pull_part_device     decode side: copies directory entries first_block .. first_block + part_blocks and the
 *                             payload they span out of d_stream into d_part (laid out for flic_split_finish_device);
 *                             *d_part_bytes (device) = size of the part stream.  The stream is treated as input: its
 *                             header and the two directory entries are validated against stream_bytes (the bytes
 *                             readable at d_stream) before anything is copied (FLIC_E_FORMAT through flic_check). */
int flic_encode_plan_device(flic_ctx *ctx, const uint8_t *d_pixels, uint32_t w, uint32_t h, uint32_t c,
                            uint32_t flags, uint64_t *d_payload_words, void *stream);
int flic_encode_emit_device(flic_ctx *ctx, uint8_t *d_stream, uint64_t capacity_bytes, uint32_t total_blocks,
                            uint32_t first_block, const uint64_t *d_base_words, void *stream);
int flic_splice_header_device(flic_ctx *ctx, uint8_t *d_stream, uint64_t capacity_bytes, uint32_t w,
                              uint32_t h_total, uint32_t c, uint32_t flags, const uint64_t *d_total_words,
                              void *stream);
int flic_pull_part_device(flic_ctx *ctx, const uint8_t *d_stream, uint64_t stream_bytes, uint32_t total_blocks,
                          uint32_t first_block, uint32_t part_blocks, uint8_t *d_part, uint64_t capacity_bytes,
                          uint64_t *d_part_bytes, void *stream);

/* ---- stage-level entry points (used by the parity tests) ---------------- */
/* Per-block residual histograms: d_hist[n_blocks_total][256] u16, flat channels left out;
 * d_flat (may be NULL): [n_blocks_total][2] u32 = {flat-channel mask, packed flat values}. */
int flic_stage_histograms(flic_ctx *ctx, const uint8_t *d_pixels, uint32_t n, uint32_t w,
                          uint32_t h, uint32_t c, uint32_t flags, uint16_t *d_hist,
                          uint32_t *d_flat, void *stream);
/* Per-block code tables from histograms: d_table[n_blocks_total][256] u16,
 * entry = len << 12 | code (len 15 = sole symbol); d_bits (may be NULL): [n_blocks_total] u32 =
 * sum over symbols of count x code length (what fixes a block's slot size). */
int flic_stage_tables(flic_ctx *ctx, const uint16_t *d_hist, uint64_t n_blocks_total,
                      uint16_t *d_table, uint32_t *d_bits, void *stream);

/* ---- measurement hooks (bench.py) --------------------------------------- */
#define FLIC_K_HISTOGRAMS 0
#define FLIC_K_TABLES 1
#define FLIC_K_PACK 2
#define FLIC_K_FINALIZE 3
#define FLIC_K_DECODE 4
#define FLIC_K_SLOTS 5
#define FLIC_K_ENCODE 6      /* the fused single-pass encoder */
#define FLIC_K_DECODE_ONE 7  /* the self-synchronising decoder of FLIC_FLAG_ONE_STREAM streams */
#define FLIC_K_COUNT 8
/* When enabled, every kernel launched through ctx is bracketed by CUDA events
 * recorded on the launching stream.  flic_get_kernel_times() waits for the
 * recorded events, returns summed milliseconds and launch counts per kernel
 * since the previous call, and clears them. */
int flic_set_kernel_timing(flic_ctx *ctx, int enable);
int flic_get_kernel_times(flic_ctx *ctx, double ms[FLIC_K_COUNT], uint64_t counts[FLIC_K_COUNT]);

/* Debug: when the context was created with FLIC_PHASE_CLOCKS=1 in the environment, thread 0 of every CTA of the
 * two block-pipelined kernels accumulates the SM cycles it spent in each phase of a block; this reads and clears
 * the sums.  cycles[0..7]: k_encode (0 ticket+clear, 1 load+residuals+histogram, 2 histogram reduce, 3 code table,
 * 4 pack, 5 look-back, 6 copy-out); cycles[8..15]: k_decode_one (0 stream copy+LUT, 1 speculative chains,
 * 2 correction rounds, 3 entry-offset table, 4 walk, 5 scan, 6 final decode, 7 un-prediction+stores).
 * FLIC_E_UNSUPPORTED otherwise (the kernels then carry a null pointer and measure nothing). */
#define FLIC_PHASES 8
int flic_get_phase_clocks(flic_ctx *ctx, uint64_t cycles[2 * FLIC_PHASES]);

/* Number of kernel launches issued through ctx since creation (bench.py's gpu_launches). */
uint64_t flic_launch_count(const flic_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
