/*
 * flp0_oracle.h — CPU model of the PROVISIONAL bitstream "FLP0".
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing outside tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may include, link or
 * call this.  The product path (fast-losless-image-compression-format_b200/)
 * never does.
 *
 * PARITY STATUS: **unpinned (licensing gate)**.  This is NOT a restatement of
 * wouter-rombouts/fast-losless-image-compression-format.  That repository's
 * source was not read (see LICENSING.md); no file:line citations into it exist
 * because nothing here follows it.  FLP0 is a block-predictive canonical-
 * Huffman format designed in this repo from textbook components so that every
 * stage BASELINE.json's north_star names (block prediction + residuals,
 * per-block histogram, canonical Huffman build, bit packing, LUT decode with
 * per-block offset tables) has a bit-exact CPU model to test the CUDA engine
 * against.  The format is specified in DESIGN.md §"FLP0".
 */
#ifndef FLP0_ORACLE_H
#define FLP0_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLP0_MAGIC 0x30504C46u /* 'F','L','P','0' little-endian */
#define FLP0_VERSION 3
#define FLP0_MAX_CODE_LEN 10
#define FLP0_LEN_SOLE 15 /* nibble value: the block's only symbol, zero-length code */
#define FLP0_HEADER_BYTES 32

#define FLP0_PRED_LEFT 1        /* x>0: left; x==0,y>0: up; (0,0): 0 — all within the block */
#define FLP0_FLAG_PRED_MASK 0x0F
#define FLP0_FLAG_SUBGREEN 0x10 /* R-=G, B-=G (mod 256) before prediction; channels >= 3 */
/* DESIGN.md §FLP0.8 (round 2): two optional layouts, recorded in the header's flags byte.
 * ONE_STREAM: the block header is 32 nibble words + 2 flat words, followed by ONE bit stream holding
 *   all of the block's symbols in row-major order, zero-padded to a word once; no row word counts, no
 *   row interleave, no slack.  EXACT: row sub-streams as in §FLP0.5-6, but the block occupies exactly
 *   the words it uses (no slot slack, §FLP0.7 does not apply).  The two cannot be combined. */
#define FLP0_FLAG_ONE_STREAM 0x20
#define FLP0_FLAG_EXACT 0x40
#define FLP0_FLAGS_ALL 0x7F

/* error codes (negative returns) */
#define FLP0_E_ARG (-1)
#define FLP0_E_CAPACITY (-2)
#define FLP0_E_FORMAT (-3)

/* Upper bound on the encoded size in bytes of one w x h x c image. */
size_t flp0_max_stream_bytes(uint32_t w, uint32_t h, uint32_t c, uint32_t bw, uint32_t bh);

/* Encode one interleaved 8-bit image. Returns stream bytes, or a negative error. */
int64_t flp0_encode(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                    uint32_t bw, uint32_t bh, uint8_t *out, size_t out_capacity);

/* Parse the 32-byte header. Returns 0 or a negative error. */
int flp0_peek(const uint8_t *stream, size_t size, uint32_t *w, uint32_t *h, uint32_t *c,
              uint32_t *flags, uint32_t *bw, uint32_t *bh, uint32_t *n_blocks,
              uint32_t *payload_words);

/* Decode one stream into w*h*c bytes. Returns 0 or a negative error. */
int flp0_decode(const uint8_t *stream, size_t size, uint8_t *pixels, size_t pixels_capacity);

/* Stage-level entry points, exposed so tests can compare the CUDA stages one by one. */

/* Residual bytes of one block, row-major [bha][bwa*c]; returns count written. */
size_t flp0_block_residuals(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t c,
                            uint32_t flags, uint32_t x0, uint32_t y0, uint32_t bw, uint32_t bh,
                            uint8_t *res);

/* Residual histogram of one block, flat channels excluded; returns the
 * flat-channel mask and writes the flat channels' values (0 for the others). */
uint32_t flp0_block_histogram(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                              uint32_t x0, uint32_t y0, uint32_t bw, uint32_t bh, uint32_t hist[256],
                              uint8_t flat_val[4]);

/* Length-limited Huffman code lengths from a 256-bin histogram. */
void flp0_build_lengths(const uint32_t hist[256], uint8_t len[256]);

/* Canonical MSB-first codes from lengths (len 0 / 15 get code 0). */
void flp0_assign_codes(const uint8_t len[256], uint16_t code[256]);

#ifdef __cplusplus
}
#endif
#endif
