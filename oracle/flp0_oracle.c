/*
 * flp0_oracle.c — scalar CPU model of the provisional FLP0 bitstream.
 *
 * TEST INFRASTRUCTURE ONLY; see flp0_oracle.h for the rules and for the
 * PARITY STATUS: unpinned (licensing gate) — this does not follow, and was
 * written without reading, the reference's source.
 *
 * Every function is a direct, deliberately simple statement of DESIGN.md
 * §FLP0 so that it can serve as the byte-exact checker for the CUDA engine.
 */
#include "flp0_oracle.h"

#include <stdlib.h>
#include <string.h>

#define L FLP0_MAX_CODE_LEN

static inline void put_u32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static inline void put_u16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static inline uint32_t get_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint32_t get_u16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

static inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

static inline uint8_t xform(const uint8_t *px, uint32_t c, uint32_t ch, uint32_t flags);

/* per-block payload bound = 32 length words + bh/2 row-count
 * words + 2 flat-channel words + bh rows of ceil(bw*c*L/32) words. */
#define FLATW 2u
/* a block occupies a SLOT whose size follows from its histogram and code
 * lengths alone — header + floor(total code bits / 32) + one word per real row (none if there
 * are no code bits at all) — so that every
 * block's position is known before any bit is packed (no serial dependence between blocks in the
 * encoder).  Rows are word-aligned, so the payload needs at most that; the rest of the slot is
 * zero. */
static size_t block_max_words(uint32_t c, uint32_t bw, uint32_t bh) {
    return 32u + bh / 2u + FLATW + (size_t)bh * (ceil_div(bw * c * L, 32u) + 1u);
}

/* a channel is FLAT in a block when every pixel of the block has the same
 * (colour-transformed) value in it — e.g. an opaque alpha plane.  Flat channels are named in the
 * block header with their value and contribute no symbols: not to the histogram, not to the row
 * streams.  The rule is mandatory (a flat channel MUST be marked), which keeps streams unique. */
static uint32_t block_flat_channels(const uint8_t *pixels, uint32_t w, uint32_t c, uint32_t flags, uint32_t x0,
                                    uint32_t y0, uint32_t bwa, uint32_t bha, uint8_t val[4]) {
    uint32_t mask = (1u << c) - 1u;
    const uint8_t *p0 = pixels + ((size_t)y0 * w + x0) * c;
    for (uint32_t ch = 0; ch < 4; ++ch) val[ch] = ch < c ? xform(p0, c, ch, flags) : 0;
    for (uint32_t y = 0; y < bha && mask; ++y)
        for (uint32_t x = 0; x < bwa; ++x) {
            const uint8_t *p = pixels + ((size_t)(y0 + y) * w + (x0 + x)) * c;
            for (uint32_t ch = 0; ch < c; ++ch)
                if (xform(p, c, ch, flags) != val[ch]) mask &= ~(1u << ch);
        }
    for (uint32_t ch = 0; ch < 4; ++ch)
        if (!((mask >> ch) & 1u)) val[ch] = 0;
    return mask;
}

size_t flp0_max_stream_bytes(uint32_t w, uint32_t h, uint32_t c, uint32_t bw, uint32_t bh) {
    size_t nb = (size_t)ceil_div(w, bw) * ceil_div(h, bh);
    return FLP0_HEADER_BYTES + 4 * (nb + 1) + 4 * nb * block_max_words(c, bw, bh);
}

/* colour transform then in-block prediction. */
static inline uint8_t xform(const uint8_t *px, uint32_t c, uint32_t ch, uint32_t flags) {
    uint8_t v = px[ch];
    if ((flags & FLP0_FLAG_SUBGREEN) && c >= 3 && (ch == 0 || ch == 2)) v = (uint8_t)(v - px[1]);
    return v;
}

size_t flp0_block_residuals(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t c,
                            uint32_t flags, uint32_t x0, uint32_t y0, uint32_t bw, uint32_t bh,
                            uint8_t *res) {
    uint32_t bwa = (w - x0 < bw) ? w - x0 : bw;
    uint32_t bha = (h - y0 < bh) ? h - y0 : bh;
    size_t n = 0;
    for (uint32_t y = 0; y < bha; ++y) {
        for (uint32_t x = 0; x < bwa; ++x) {
            const uint8_t *p = pixels + ((size_t)(y0 + y) * w + (x0 + x)) * c;
            for (uint32_t ch = 0; ch < c; ++ch) {
                uint8_t v = xform(p, c, ch, flags), pred;
                if (x > 0) pred = xform(p - c, c, ch, flags);
                else if (y > 0) pred = xform(p - (size_t)w * c, c, ch, flags);
                else pred = 0;
                res[n++] = (uint8_t)(v - pred);
            }
        }
    }
    return n;
}

/* Residual histogram of one block with its flat channels left out (what k_histograms emits). */
uint32_t flp0_block_histogram(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                              uint32_t x0, uint32_t y0, uint32_t bw, uint32_t bh, uint32_t hist[256],
                              uint8_t flat_val[4]) {
    uint32_t bwa = (w - x0 < bw) ? w - x0 : bw, bha = (h - y0 < bh) ? h - y0 : bh;
    uint8_t *res = (uint8_t *)malloc((size_t)bw * bh * c);
    memset(hist, 0, 256 * sizeof(uint32_t));
    if (!res) return 0;
    size_t n = flp0_block_residuals(pixels, w, h, c, flags, x0, y0, bw, bh, res);
    uint32_t flat = block_flat_channels(pixels, w, c, flags, x0, y0, bwa, bha, flat_val);
    for (size_t i = 0; i < n; ++i)
        if (!((flat >> (i % c)) & 1u)) hist[res[i]]++;
    free(res);
    return flat;
}

/* (1) sort active symbols by (count, symbol) ascending;
 * (2) two-queue Huffman merge, a leaf wins a tie against an internal node;
 * (3) count leaves per depth, folding depths > L into L;
 * (4) Kraft repair: while over budget, drop one L-code and split the deepest
 *     shorter code (the classic deflate-encoder fix-up);
 * (5) hand lengths out by sorted rank, rarest symbols get the longest codes. */
void flp0_build_lengths(const uint32_t hist[256], uint8_t len[256]) {
    uint16_t order[256];
    uint32_t lw[256], iw[256];
    uint16_t lpar[256], ipar[256], idep[256];
    uint32_t num[L + 1];
    int n = 0;

    memset(len, 0, 256);
    for (int s = 0; s < 256; ++s)
        if (hist[s]) order[n++] = (uint16_t)s;
    if (n == 0) return;
    if (n == 1) { len[order[0]] = FLP0_LEN_SOLE; return; }

    /* insertion sort by (count, symbol); symbols already ascending so it is stable */
    for (int i = 1; i < n; ++i) {
        uint16_t s = order[i];
        int j = i - 1;
        while (j >= 0 && hist[order[j]] > hist[s]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = s;
    }
    for (int i = 0; i < n; ++i) lw[i] = hist[order[i]];

    int li = 0, ii = 0;
    for (int k = 0; k < n - 1; ++k) {
        uint32_t wsum = 0;
        for (int t = 0; t < 2; ++t) {
            if (li < n && (ii >= k || lw[li] <= iw[ii])) { wsum += lw[li]; lpar[li++] = (uint16_t)k; }
            else { wsum += iw[ii]; ipar[ii++] = (uint16_t)k; }
        }
        iw[k] = wsum;
    }
    idep[n - 2] = 0;
    for (int k = n - 3; k >= 0; --k) idep[k] = (uint16_t)(idep[ipar[k]] + 1);

    memset(num, 0, sizeof num);
    for (int i = 0; i < n; ++i) {
        uint32_t d = idep[lpar[i]] + 1u;
        num[d > L ? L : d]++;
    }
    uint32_t total = 0;
    for (int l = L; l >= 1; --l) total += num[l] << (L - l);
    while (total > (1u << L)) {
        num[L]--;
        for (int l = L - 1; l >= 1; --l)
            if (num[l]) { num[l]--; num[l + 1] += 2; break; }
        total--;
    }
    int idx = 0;
    for (int l = L; l >= 1; --l)
        for (uint32_t j = 0; j < num[l]; ++j) len[order[idx++]] = (uint8_t)l;
}

/* canonical codes, MSB-first, ordered by (length, symbol). */
void flp0_assign_codes(const uint8_t len[256], uint16_t code[256]) {
    uint32_t num[L + 2], next[L + 2];
    memset(num, 0, sizeof num);
    for (int s = 0; s < 256; ++s)
        if (len[s] >= 1 && len[s] <= L) num[len[s]]++;
    next[0] = 0; next[1] = 0;
    for (int l = 2; l <= L; ++l) next[l] = (next[l - 1] + num[l - 1]) << 1;
    for (int s = 0; s < 256; ++s) {
        if (len[s] >= 1 && len[s] <= L) code[s] = (uint16_t)next[len[s]]++;
        else code[s] = 0;
    }
}

/* MSB-first bit writer over little-endian u32 words. */
typedef struct { uint8_t *p; uint64_t acc; int nacc; uint32_t words; } bitw;
static inline void bw_put(bitw *b, uint32_t code, int n) {
    b->acc = (b->acc << n) | code; b->nacc += n;
    if (b->nacc >= 32) {
        b->nacc -= 32;
        put_u32(b->p, (uint32_t)(b->acc >> b->nacc));
        b->p += 4; b->words++;
    }
}
static inline void bw_flush(bitw *b) {
    if (b->nacc > 0) {
        put_u32(b->p, (uint32_t)(b->acc << (32 - b->nacc)));
        b->p += 4; b->words++; b->nacc = 0;
    }
    b->acc = 0;
}

int64_t flp0_encode(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t c, uint32_t flags,
                    uint32_t bw, uint32_t bh, uint8_t *out, size_t out_capacity) {
    if (!pixels || !out || w == 0 || h == 0 || c < 1 || c > 4 || bw == 0 || bh == 0 || (bh & 1) ||
        (flags & FLP0_FLAG_PRED_MASK) != FLP0_PRED_LEFT || (flags & ~(uint32_t)FLP0_FLAGS_ALL) || bw * c * L / 32 > 65535u ||
        ((flags & FLP0_FLAG_ONE_STREAM) && (flags & FLP0_FLAG_EXACT)))
        return FLP0_E_ARG;
    if (out_capacity < flp0_max_stream_bytes(w, h, c, bw, bh)) return FLP0_E_CAPACITY;

    uint32_t nbx = ceil_div(w, bw), nby = ceil_div(h, bh), nb = nbx * nby;
    uint8_t *dir = out + FLP0_HEADER_BYTES;
    uint8_t *payload = dir + 4 * ((size_t)nb + 1);
    if (bh > 32768) return FLP0_E_ARG;
    uint8_t *res = (uint8_t *)malloc((size_t)bw * bh * c);
    uint8_t *rowbuf = (uint8_t *)malloc(4 * (size_t)bh * ceil_div(bw * c * L, 32u) + 4);
    if (!res || !rowbuf) { free(res); free(rowbuf); return FLP0_E_ARG; }

    uint32_t wpos = 0; /* payload position in words */
    for (uint32_t by = 0; by < nby; ++by) {
        for (uint32_t bx = 0; bx < nbx; ++bx) {
            uint32_t x0 = bx * bw, y0 = by * bh;
            uint32_t bwa = (w - x0 < bw) ? w - x0 : bw, bha = (h - y0 < bh) ? h - y0 : bh;
            uint32_t rowsym = bwa * c;
            uint32_t hist[256];
            uint8_t len[256];
            uint16_t code[256];
            size_t n = flp0_block_residuals(pixels, w, h, c, flags, x0, y0, bw, bh, res);
            uint8_t fval[4];
            uint32_t flat = block_flat_channels(pixels, w, c, flags, x0, y0, bwa, bha, fval);

            memset(hist, 0, sizeof hist);
            for (size_t i = 0; i < n; ++i)
                if (!((flat >> (i % c)) & 1u)) hist[res[i]]++;
            flp0_build_lengths(hist, len);
            flp0_assign_codes(len, code);

            put_u32(dir + 4 * (size_t)(by * nbx + bx), wpos);
            uint8_t *blk = payload + 4 * (size_t)wpos;
            for (int s = 0; s < 256; s += 2) blk[s >> 1] = (uint8_t)(len[s] | (len[s + 1] << 4));
            if (flags & FLP0_FLAG_ONE_STREAM) {
                /* flat words right after the nibbles, then one bit stream for the whole block */
                uint8_t *fw1 = blk + 128;
                put_u32(fw1, flat);
                fw1[4] = fval[0]; fw1[5] = fval[1]; fw1[6] = fval[2]; fw1[7] = fval[3];
                bitw b = { fw1 + 4 * FLATW, 0, 0, 0 };
                for (uint32_t y = 0; y < bha; ++y) {
                    const uint8_t *r = res + (size_t)y * rowsym;
                    for (uint32_t i = 0; i < rowsym; ++i) {
                        uint8_t l = len[r[i]];
                        if ((flat >> (i % c)) & 1u) continue;
                        if (l != FLP0_LEN_SOLE) bw_put(&b, code[r[i]], l);
                    }
                }
                bw_flush(&b);
                wpos += 32u + FLATW + b.words;
                continue;
            }
            uint8_t *rw = blk + 128;
            /* each row is bit-packed on its own ... */
            static _Thread_local uint32_t rwc[32768];
            uint32_t rowcap = ceil_div(bw * c * L, 32u), minw = 0xFFFFFFFFu, total = 0;
            for (uint32_t y = 0; y < bh; ++y) {
                bitw b = { rowbuf + 4 * (size_t)y * rowcap, 0, 0, 0 };
                if (y < bha) {
                    const uint8_t *r = res + (size_t)y * rowsym;
                    for (uint32_t i = 0; i < rowsym; ++i) {
                        uint8_t l = len[r[i]];
                        if ((flat >> (i % c)) & 1u) continue; /* flat channel: no symbol */
                        if (l != FLP0_LEN_SOLE) bw_put(&b, code[r[i]], l);
                    }
                    bw_flush(&b);
                    if (b.words < minw) minw = b.words;
                }
                rwc[y] = b.words;
                total += b.words;
                put_u16(rw + 2 * y, b.words);
            }
            /* ... §FLP0.6: then the first minw words of the bha real rows are interleaved
             * (word k of row r at k*bha + r) and the rows' tails follow in row order. */
            uint64_t code_bits = 0;
            for (int sy = 0; sy < 256; ++sy)
                if (len[sy] >= 1 && len[sy] <= L) code_bits += (uint64_t)hist[sy] * len[sy];
            uint32_t slot = 32u + bh / 2u + FLATW + (uint32_t)(code_bits >> 5) + (code_bits ? bha : 0u);
            if (flags & FLP0_FLAG_EXACT) slot = 32u + bh / 2u + FLATW + total; /* §FLP0.8: exactly the words used */
            uint8_t *fw = rw + 2 * bh; /* flat-channel words: mask, then the four values */
            put_u32(fw, flat);
            fw[4] = fval[0]; fw[5] = fval[1]; fw[6] = fval[2]; fw[7] = fval[3];
            uint8_t *o = fw + 4 * FLATW;
            for (uint32_t k = 0; k < minw; ++k)
                for (uint32_t y = 0; y < bha; ++y, o += 4) memcpy(o, rowbuf + 4 * ((size_t)y * rowcap + k), 4);
            for (uint32_t y = 0; y < bha; ++y) {
                memcpy(o, rowbuf + 4 * ((size_t)y * rowcap + minw), 4 * (size_t)(rwc[y] - minw));
                o += 4 * (size_t)(rwc[y] - minw);
            }
            /* zero the slack up to the block's slot */
            memset(o, 0, 4 * (size_t)(slot - (32u + bh / 2u + FLATW + total)));
            wpos += slot;
        }
    }
    free(rowbuf);
    free(res);
    put_u32(dir + 4 * (size_t)nb, wpos);

    put_u32(out + 0, FLP0_MAGIC);
    put_u16(out + 4, FLP0_VERSION);
    out[6] = (uint8_t)c; out[7] = (uint8_t)flags;
    put_u32(out + 8, w); put_u32(out + 12, h);
    put_u16(out + 16, bw); put_u16(out + 18, bh);
    put_u32(out + 20, nb); put_u32(out + 24, wpos); put_u32(out + 28, L);
    return (int64_t)(FLP0_HEADER_BYTES + 4 * ((size_t)nb + 1) + 4 * (size_t)wpos);
}

int flp0_peek(const uint8_t *s, size_t size, uint32_t *w, uint32_t *h, uint32_t *c, uint32_t *flags,
              uint32_t *bw, uint32_t *bh, uint32_t *n_blocks, uint32_t *payload_words) {
    if (!s || size < FLP0_HEADER_BYTES) return FLP0_E_FORMAT;
    if (get_u32(s) != FLP0_MAGIC || get_u16(s + 4) != FLP0_VERSION || get_u32(s + 28) != L)
        return FLP0_E_FORMAT;
    uint32_t W = get_u32(s + 8), H = get_u32(s + 12), C = s[6], F = s[7];
    uint32_t BW = get_u16(s + 16), BH = get_u16(s + 18), NB = get_u32(s + 20), PW = get_u32(s + 24);
    if (W == 0 || H == 0 || C < 1 || C > 4 || BW == 0 || BH == 0 || (BH & 1) ||
        (F & FLP0_FLAG_PRED_MASK) != FLP0_PRED_LEFT || (F & ~(uint32_t)FLP0_FLAGS_ALL) ||
        ((F & FLP0_FLAG_ONE_STREAM) && (F & FLP0_FLAG_EXACT)))
        return FLP0_E_FORMAT;
    if ((uint64_t)ceil_div(W, BW) * ceil_div(H, BH) != NB) return FLP0_E_FORMAT;
    if ((uint64_t)FLP0_HEADER_BYTES + 4 * ((uint64_t)NB + 1) + 4 * (uint64_t)PW > size) return FLP0_E_FORMAT;
    if (w) *w = W;
    if (h) *h = H;
    if (c) *c = C;
    if (flags) *flags = F;
    if (bw) *bw = BW;
    if (bh) *bh = BH;
    if (n_blocks) *n_blocks = NB;
    if (payload_words) *payload_words = PW;
    return 0;
}

int flp0_decode(const uint8_t *s, size_t size, uint8_t *pixels, size_t cap) {
    uint32_t w, h, c, flags, bw, bh, nb, pw;
    int rc = flp0_peek(s, size, &w, &h, &c, &flags, &bw, &bh, &nb, &pw);
    if (rc) return rc;
    if (!pixels || cap < (size_t)w * h * c) return FLP0_E_CAPACITY;
    const uint8_t *dir = s + FLP0_HEADER_BYTES;
    const uint8_t *payload = dir + 4 * ((size_t)nb + 1);
    uint32_t nbx = ceil_div(w, bw);
    uint16_t *lut = (uint16_t *)malloc(sizeof(uint16_t) << L);
    if (!lut) return FLP0_E_ARG;

    for (uint32_t b = 0; b < nb; ++b) {
        uint32_t off = get_u32(dir + 4 * (size_t)b), end = get_u32(dir + 4 * (size_t)(b + 1));
        const int one = (flags & FLP0_FLAG_ONE_STREAM) != 0;
        if (off > end || end > pw || end - off < (one ? 32u + FLATW : 32u + bh / 2u + FLATW)) { free(lut); return FLP0_E_FORMAT; }
        const uint8_t *blk = payload + 4 * (size_t)off;
        uint32_t x0 = (b % nbx) * bw, y0 = (b / nbx) * bh;
        uint32_t bwa = (w - x0 < bw) ? w - x0 : bw, bha = (h - y0 < bh) ? h - y0 : bh;
        uint8_t len[256];
        uint16_t code[256];
        int sole = -1;
        for (int sy = 0; sy < 256; sy += 2) { len[sy] = blk[sy >> 1] & 15; len[sy + 1] = blk[sy >> 1] >> 4; }
        for (int sy = 0; sy < 256; ++sy) {
            if (len[sy] == FLP0_LEN_SOLE) sole = sy;
            else if (len[sy] > L) { free(lut); return FLP0_E_FORMAT; }
        }
        flp0_assign_codes(len, code);
        /* LUT indexed by the next L bits (MSB-first): entry = sym | len << 8; 0xFFFF = invalid */
        memset(lut, 0xFF, sizeof(uint16_t) << L);
        for (int sy = 0; sy < 256; ++sy) {
            if (len[sy] >= 1 && len[sy] <= L) {
                uint32_t first = (uint32_t)code[sy] << (L - len[sy]), cnt = 1u << (L - len[sy]);
                if (first + cnt > (1u << L)) { free(lut); return FLP0_E_FORMAT; }
                for (uint32_t i = 0; i < cnt; ++i) lut[first + i] = (uint16_t)(sy | (len[sy] << 8));
            }
        }
        const uint8_t *rw = blk + 128;
        const uint8_t *fw = one ? blk + 128 : rw + 2 * bh;
        uint32_t flat = get_u32(fw);
        if (flat >> c) { free(lut); return FLP0_E_FORMAT; }
        const uint8_t *body = fw + 4 * FLATW;
        uint32_t minw = 0xFFFFFFFFu, total = 0;
        /* §FLP0.8: one stream for the whole block — the reader state below carries over from row to row,
         * and words past the block's extent read as zeros */
        uint64_t acc1 = 0; int nacc1 = 0; uint32_t used1 = 0;
        const uint32_t words1 = end - off - (32u + FLATW);
        for (uint32_t y = 0; y < bh && !one; ++y) {
            uint32_t words = get_u16(rw + 2 * y);
            if (y >= bha && words) { free(lut); return FLP0_E_FORMAT; }
            if (y < bha && words < minw) minw = words;
            total += words;
        }
        if (!one && 32u + bh / 2u + FLATW + total > end - off) { free(lut); return FLP0_E_FORMAT; }
        if (one) minw = 0;
        const uint8_t *tail = body + 4 * (size_t)minw * bha;
        for (uint32_t y = 0; y < bha; ++y) {
            uint32_t words = one ? words1 : get_u16(rw + 2 * y);
            {
                uint8_t *dst = pixels + ((size_t)(y0 + y) * w + x0) * c;
                const uint8_t *up = dst - (size_t)w * c;
                uint64_t acc = 0; int nacc = 0; uint32_t used = 0;
                if (one) { acc = acc1; nacc = nacc1; used = used1; }
                uint8_t prev[4] = {0, 0, 0, 0};
                for (uint32_t x = 0; x < bwa; ++x) {
                    uint8_t t[4];
                    for (uint32_t ch = 0; ch < c; ++ch) {
                        uint8_t r;
                        if ((flat >> ch) & 1u) { t[ch] = fw[4 + ch]; continue; }
                        if (sole >= 0) r = (uint8_t)sole;
                        else {
                            if (nacc < L) {
                                uint32_t wv = 0;
                                if (used < minw) wv = get_u32(body + 4 * ((size_t)used * bha + y));
                                else if (used < words) wv = get_u32(tail + 4 * (size_t)(used - minw));
                                used++;
                                acc = (acc << 32) | wv; nacc += 32;
                            }
                            uint16_t e = lut[(acc >> (nacc - L)) & ((1u << L) - 1)];
                            if (e == 0xFFFF) { free(lut); return FLP0_E_FORMAT; }
                            r = (uint8_t)e; nacc -= e >> 8;
                        }
                        uint8_t pred;
                        if (x > 0) pred = prev[ch];
                        else if (y > 0) {
                            pred = up[ch];
                            if ((flags & FLP0_FLAG_SUBGREEN) && c >= 3 && (ch == 0 || ch == 2))
                                pred = (uint8_t)(pred - up[1]);
                        } else pred = 0;
                        t[ch] = (uint8_t)(r + pred);
                    }
                    for (uint32_t ch = 0; ch < c; ++ch) prev[ch] = t[ch];
                    if ((flags & FLP0_FLAG_SUBGREEN) && c >= 3) { t[0] = (uint8_t)(t[0] + t[1]); t[2] = (uint8_t)(t[2] + t[1]); }
                    for (uint32_t ch = 0; ch < c; ++ch) dst[x * c + ch] = t[ch];
                }
                if (one) { acc1 = acc; nacc1 = nacc; used1 = used; }
            }
            if (!one) tail += 4 * (size_t)(words - minw);
        }
    }
    free(lut);
    return 0;
}
