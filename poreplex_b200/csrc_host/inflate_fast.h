// inflate_fast.h -- zlib-format (RFC 1950 / 1951) decompressor for the deflate-filtered chunks
// of FAST5 files.  One call decodes one whole stream (HDF5 chunks are small, complete zlib
// streams), so the decoder is organised for that case instead of zlib's resumable state machine:
//   * 64-bit bit buffer refilled with one unaligned load,
//   * table-driven Huffman decoding (10-bit root for literals/lengths, 8-bit root for
//     distances, second-level tables for longer codes) -- one lookup per symbol,
//   * a fast loop that runs while >= 16 input bytes and >= 272 output bytes remain, so it needs
//     no bounds checks (up to three literals or one match per iteration), and a careful loop
//     for the tail,
//   * matches copied eight bytes at a time.
// Errors (truncated / corrupt input, output overflow, Adler-32 mismatch) are negative return
// values; nothing is written past `out + cap` and nothing is read past `in + n`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace pbinf {

enum : int64_t {
    ERR_TRUNCATED = -1, ERR_CORRUPT = -2, ERR_OVERFLOW = -3, ERR_HEADER = -4, ERR_CHECKSUM = -5,
};

// table entry: bits 0-7 code length to consume, 8-11 kind, 12-15 extra-bit count (or sub-table
// index width), 16-31 value (literal, base of a length / distance, or sub-table offset)
enum : uint32_t { K_LITERAL = 0, K_BASE = 1, K_END = 2, K_SUBTABLE = 3, K_INVALID = 4 };
constexpr int LIT_ROOT = 10, DIST_ROOT = 8;
constexpr int LIT_TABLE = (1 << LIT_ROOT) + 2048, DIST_TABLE = (1 << DIST_ROOT) + 1024;

struct Tables {
    uint32_t lit[LIT_TABLE];
    uint32_t dist[DIST_TABLE];
    uint32_t pre[128];
    bool fixed_ready = false;
    uint32_t fixed_lit[LIT_TABLE];
    uint32_t fixed_dist[DIST_TABLE];
};

namespace detail {

inline uint32_t entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t len)
{
    return (value << 16) | (extra << 12) | (kind << 8) | len;
}

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35,
                                      43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3,
                                      4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193,
                                       257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                       8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8,
                                       9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// what a decoded symbol means, as a table entry without its code length
inline uint32_t litlen_symbol(unsigned sym)
{
    if (sym < 256) return entry(sym, K_LITERAL, 0, 0);
    if (sym == 256) return entry(0, K_END, 0, 0);
    if (sym < 286) return entry(LEN_BASE[sym - 257], K_BASE, LEN_EXTRA[sym - 257], 0);
    return entry(0, K_INVALID, 0, 0);
}
inline uint32_t dist_symbol(unsigned sym)
{
    if (sym < 30) return entry(DIST_BASE[sym], K_BASE, DIST_EXTRA[sym], 0);
    return entry(0, K_INVALID, 0, 0);
}
inline uint32_t pre_symbol(unsigned sym) { return entry(sym, K_LITERAL, 0, 0); }

inline unsigned reverse_bits(unsigned code, int len)       // len <= 15
{
    unsigned v = code;
    v = ((v & 0x5555u) << 1) | ((v >> 1) & 0x5555u);
    v = ((v & 0x3333u) << 2) | ((v >> 2) & 0x3333u);
    v = ((v & 0x0F0Fu) << 4) | ((v >> 4) & 0x0F0Fu);
    v = ((v & 0x00FFu) << 8) | ((v >> 8) & 0x00FFu);
    return v >> (16 - len);
}

// Canonical Huffman decoding table from code lengths.  Returns false for an over-subscribed
// code or an incomplete one; like zlib, a literal/length or distance code (`allow_incomplete`)
// may consist of a single 1-bit symbol or, for distances, of nothing at all -- unassigned slots
// then decode to K_INVALID.
template <class SymbolFn>
bool build_table(const uint8_t *lens, int n_sym, int root, uint32_t *table, int table_cap,
                 SymbolFn symbol, bool allow_incomplete)
{
    int count[16] = {0};
    for (int i = 0; i < n_sym; i++) count[lens[i]]++;
    count[0] = 0;
    int max_len = 15;
    while (max_len > 0 && count[max_len] == 0) max_len--;
    for (int i = 0; i < (1 << root); i++) table[i] = entry(0, K_INVALID, 0, 1);
    if (max_len == 0) return allow_incomplete;           // no codes at all
    long left = 1;
    for (int len = 1; len <= 15; len++) {
        left = (left << 1) - count[len];
        if (left < 0) return false;                      // over-subscribed
    }
    if (left > 0 && !(allow_incomplete && max_len == 1 && count[1] == 1)) return false;
    unsigned next_code[16];
    unsigned code = 0;
    for (int len = 1; len <= 15; len++) {
        code = (code + count[len - 1]) << 1;
        next_code[len] = code;
    }
    // second-level tables: one per root prefix that has codes longer than `root`; its width is
    // the longest such code minus root
    int sub_bits_of[1 << 10];                            // root <= 10
    if (max_len > root) {
        for (int i = 0; i < (1 << root); i++) sub_bits_of[i] = 0;
        unsigned nc[16];
        memcpy(nc, next_code, sizeof nc);
        for (int s = 0; s < n_sym; s++) {
            const int len = lens[s];
            if (len <= root) { if (len) nc[len]++; continue; }
            const unsigned rev = reverse_bits(nc[len]++, len);
            const int prefix = rev & ((1u << root) - 1);
            if (len - root > sub_bits_of[prefix]) sub_bits_of[prefix] = len - root;
        }
        int next_free = 1 << root;
        for (int i = 0; i < (1 << root); i++) {
            if (!sub_bits_of[i]) continue;
            const int size = 1 << sub_bits_of[i];
            if (next_free + size > table_cap) return false;
            table[i] = entry((uint32_t)next_free, K_SUBTABLE, (uint32_t)sub_bits_of[i], (uint32_t)root);
            for (int k = 0; k < size; k++) table[next_free + k] = entry(0, K_INVALID, 0, (uint32_t)root + 1);
            next_free += size;
        }
    }
    for (int s = 0; s < n_sym; s++) {
        const int len = lens[s];
        if (!len) continue;
        const unsigned rev = reverse_bits(next_code[len]++, len);
        const uint32_t e = symbol((unsigned)s) | (uint32_t)len;
        if (len <= root) {
            for (unsigned i = rev; i < (1u << root); i += 1u << len) table[i] = e;
        } else {
            const uint32_t link = table[rev & ((1u << root) - 1)];
            const unsigned base = link >> 16, sub_bits = (link >> 12) & 15;
            for (unsigned i = rev >> root; i < (1u << sub_bits); i += 1u << (len - root)) table[base + i] = e;
        }
    }
    return true;
}

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline void store64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

inline uint32_t adler32(const uint8_t *p, size_t n)
{
    uint32_t a = 1, b = 0;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        n -= k;
        while (k >= 8) {
            a += p[0]; b += a; a += p[1]; b += a; a += p[2]; b += a; a += p[3]; b += a;
            a += p[4]; b += a; a += p[5]; b += a; a += p[6]; b += a; a += p[7]; b += a;
            p += 8; k -= 8;
        }
        while (k--) { a += *p++; b += a; }
        a %= 65521; b %= 65521;
    }
    return (b << 16) | a;
}

struct Stream {
    const uint8_t *in, *in_end;
    uint8_t *out, *out_start, *out_end;
    uint64_t bitbuf = 0;
    unsigned bitcnt = 0;

    // careful refill: never past in_end (one 8-byte load while that much input remains)
    void refill_safe() {
        if (in_end - in >= 8) { refill_fast(); return; }
        while (bitcnt <= 56 && in < in_end) { bitbuf |= (uint64_t)*in++ << bitcnt; bitcnt += 8; }
    }
    // fast refill: needs in + 8 <= in_end; leaves 56..63 valid bits
    void refill_fast() {
        bitbuf |= load64(in) << bitcnt;
        const unsigned adv = (63 - bitcnt) >> 3;
        in += adv;
        bitcnt += adv * 8;
    }
    uint32_t peek(unsigned n) const { return (uint32_t)(bitbuf & ((1ull << n) - 1)); }
    void drop(unsigned n) { bitbuf >>= n; bitcnt -= n; }
    // careful "take n bits": false when the input is exhausted
    bool take(unsigned n, uint32_t &v) {
        if (bitcnt < n) { refill_safe(); if (bitcnt < n) return false; }
        v = peek(n);
        drop(n);
        return true;
    }
};

// careful decode of one symbol; <0 on error
inline int64_t decode_safe(Stream &s, const uint32_t *table, int root, uint32_t &e)
{
    if (s.bitcnt < 15) s.refill_safe();
    e = table[s.peek((unsigned)root)];
    if (((e >> 8) & 15) == K_SUBTABLE)
        e = table[(e >> 16) + ((s.bitbuf >> root) & ((1u << ((e >> 12) & 15)) - 1))];
    const unsigned len = e & 0xFF;
    if (((e >> 8) & 15) == K_INVALID) return ERR_CORRUPT;
    if (len > s.bitcnt) return ERR_TRUNCATED;
    s.drop(len);
    return 0;
}

inline int64_t copy_match(Stream &s, unsigned length, unsigned dist, bool fast)
{
    if (dist > (size_t)(s.out - s.out_start)) return ERR_CORRUPT;
    const uint8_t *src = s.out - dist;
    if (fast) {                                          // >= 266 bytes of room: may overrun by 7
        uint8_t *dst = s.out, *end = s.out + length;
        if (dist >= 8) {
            do { store64(dst, load64(src)); dst += 8; src += 8; } while (dst < end);
        } else if (dist == 1) {
            memset(dst, *src, length);
        } else {
            do { *dst++ = *src++; } while (dst < end);
        }
        s.out = end;
        return 0;
    }
    if (length > (size_t)(s.out_end - s.out)) return ERR_OVERFLOW;
    for (unsigned i = 0; i < length; i++) s.out[i] = src[i];
    s.out += length;
    return 0;
}

inline int64_t inflate_block(Stream &s, const uint32_t *lit, const uint32_t *dist)
{
    constexpr uint32_t LMASK = (1u << LIT_ROOT) - 1, DMASK = (1u << DIST_ROOT) - 1;
    // ---- fast loop ----------------------------------------------------------------------
    while (s.in_end - s.in >= 16 && s.out_end - s.out >= 272) {
        s.refill_fast();
        uint32_t e = lit[s.bitbuf & LMASK];
        if (((e >> 8) & 15) == K_SUBTABLE)
            e = lit[(e >> 16) + ((s.bitbuf >> LIT_ROOT) & ((1u << ((e >> 12) & 15)) - 1))];
        s.drop(e & 0xFF);
        if (((e >> 8) & 15) == K_LITERAL) {              // up to three literals per refill
            *s.out++ = (uint8_t)(e >> 16);
            e = lit[s.bitbuf & LMASK];
            if (((e >> 8) & 15) == K_SUBTABLE)
                e = lit[(e >> 16) + ((s.bitbuf >> LIT_ROOT) & ((1u << ((e >> 12) & 15)) - 1))];
            s.drop(e & 0xFF);
            if (((e >> 8) & 15) == K_LITERAL) {
                *s.out++ = (uint8_t)(e >> 16);
                e = lit[s.bitbuf & LMASK];
                if (((e >> 8) & 15) == K_SUBTABLE)
                    e = lit[(e >> 16) + ((s.bitbuf >> LIT_ROOT) & ((1u << ((e >> 12) & 15)) - 1))];
                s.drop(e & 0xFF);
                if (((e >> 8) & 15) == K_LITERAL) { *s.out++ = (uint8_t)(e >> 16); continue; }
            }
            s.refill_fast();                             // <= 45 bits gone: top up for the match
        }
        const uint32_t kind = (e >> 8) & 15;
        if (kind == K_END) return 0;
        if (kind != K_BASE) return ERR_CORRUPT;
        const unsigned lx = (e >> 12) & 15;
        const unsigned length = (e >> 16) + s.peek(lx);
        s.drop(lx);                                      // <= 15 + 5 bits since the last refill
        uint32_t d = dist[s.bitbuf & DMASK];
        if (((d >> 8) & 15) == K_SUBTABLE)
            d = dist[(d >> 16) + ((s.bitbuf >> DIST_ROOT) & ((1u << ((d >> 12) & 15)) - 1))];
        if (((d >> 8) & 15) != K_BASE) return ERR_CORRUPT;
        s.drop(d & 0xFF);
        const unsigned dx = (d >> 12) & 15;
        const unsigned distance = (d >> 16) + s.peek(dx);
        s.drop(dx);                                      // 20 + 15 + 13 = 48 <= 56
        const int64_t rc = copy_match(s, length, distance, true);
        if (rc) return rc;
    }
    // ---- careful loop -------------------------------------------------------------------
    for (;;) {
        uint32_t e;
        int64_t rc = decode_safe(s, lit, LIT_ROOT, e);
        if (rc) return rc;
        const uint32_t kind = (e >> 8) & 15;
        if (kind == K_LITERAL) {
            if (s.out >= s.out_end) return ERR_OVERFLOW;
            *s.out++ = (uint8_t)(e >> 16);
            continue;
        }
        if (kind == K_END) return 0;
        if (kind != K_BASE) return ERR_CORRUPT;
        uint32_t x;
        if (!s.take((e >> 12) & 15, x)) return ERR_TRUNCATED;
        const unsigned length = (e >> 16) + x;
        uint32_t d;
        rc = decode_safe(s, dist, DIST_ROOT, d);
        if (rc) return rc;
        if (((d >> 8) & 15) != K_BASE) return ERR_CORRUPT;
        if (!s.take((d >> 12) & 15, x)) return ERR_TRUNCATED;
        rc = copy_match(s, length, (d >> 16) + x, false);
        if (rc) return rc;
    }
}

inline int64_t read_dynamic_tables(Stream &s, Tables &t)
{
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t hlit, hdist, hclen, v;
    if (!s.take(5, hlit) || !s.take(5, hdist) || !s.take(4, hclen)) return ERR_TRUNCATED;
    hlit += 257; hdist += 1; hclen += 4;
    if (hlit > 286 || hdist > 30) return ERR_CORRUPT;
    uint8_t pre_lens[19] = {0};
    for (unsigned i = 0; i < hclen; i++) {
        if (!s.take(3, v)) return ERR_TRUNCATED;
        pre_lens[ORDER[i]] = (uint8_t)v;
    }
    if (!build_table(pre_lens, 19, 7, t.pre, 128, pre_symbol, false)) return ERR_CORRUPT;
    uint8_t lens[286 + 30 + 138];
    unsigned n = 0;
    while (n < hlit + hdist) {
        uint32_t e;
        const int64_t rc = decode_safe(s, t.pre, 7, e);
        if (rc) return rc;
        const unsigned sym = e >> 16;
        if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
        unsigned rep;
        uint8_t what = 0;
        if (sym == 16) {
            if (n == 0) return ERR_CORRUPT;
            what = lens[n - 1];
            if (!s.take(2, v)) return ERR_TRUNCATED;
            rep = 3 + v;
        } else if (sym == 17) {
            if (!s.take(3, v)) return ERR_TRUNCATED;
            rep = 3 + v;
        } else {
            if (!s.take(7, v)) return ERR_TRUNCATED;
            rep = 11 + v;
        }
        if (n + rep > hlit + hdist) return ERR_CORRUPT;
        memset(lens + n, what, rep);
        n += rep;
    }
    if (lens[256] == 0) return ERR_CORRUPT;              // no end-of-block code
    if (!build_table(lens, (int)hlit, LIT_ROOT, t.lit, LIT_TABLE, litlen_symbol, true)) return ERR_CORRUPT;
    if (!build_table(lens + hlit, (int)hdist, DIST_ROOT, t.dist, DIST_TABLE, dist_symbol, true)) return ERR_CORRUPT;
    return 0;
}

inline void ensure_fixed(Tables &t)
{
    if (t.fixed_ready) return;
    uint8_t lens[288];
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    build_table(lens, 288, LIT_ROOT, t.fixed_lit, LIT_TABLE, litlen_symbol, false);
    uint8_t dl[32];
    for (int i = 0; i < 32; i++) dl[i] = 5;
    build_table(dl, 32, DIST_ROOT, t.fixed_dist, DIST_TABLE, dist_symbol, false);
    t.fixed_ready = true;
}

}  // namespace detail

// raw deflate stream -> out; returns bytes written (and *consumed input bytes) or an error
inline int64_t inflate_raw(const uint8_t *in, size_t n, uint8_t *out, size_t cap, Tables &t,
                           size_t *consumed = nullptr)
{
    detail::Stream s;
    s.in = in; s.in_end = in + n;
    s.out = s.out_start = out; s.out_end = out + cap;
    for (;;) {
        uint32_t final_block, type;
        if (!s.take(1, final_block) || !s.take(2, type)) return ERR_TRUNCATED;
        if (type == 0) {                                  // stored
            s.drop(s.bitcnt & 7);                         // to the byte boundary
            uint32_t len, nlen;
            if (!s.take(16, len) || !s.take(16, nlen)) return ERR_TRUNCATED;
            if ((len ^ 0xFFFF) != nlen) return ERR_CORRUPT;
            // bytes still in the bit buffer come first
            while (len && s.bitcnt >= 8) {
                if (s.out >= s.out_end) return ERR_OVERFLOW;
                *s.out++ = (uint8_t)s.peek(8);
                s.drop(8);
                len--;
            }
            if (len) {
                s.bitbuf = 0; s.bitcnt = 0;               // whatever is left are over-read bits
                if (len > (size_t)(s.in_end - s.in)) return ERR_TRUNCATED;
                if (len > (size_t)(s.out_end - s.out)) return ERR_OVERFLOW;
                memcpy(s.out, s.in, len);
                s.out += len;
                s.in += len;
            }
        } else if (type == 1) {
            detail::ensure_fixed(t);
            const int64_t rc = detail::inflate_block(s, t.fixed_lit, t.fixed_dist);
            if (rc) return rc;
        } else if (type == 2) {
            int64_t rc = detail::read_dynamic_tables(s, t);
            if (rc) return rc;
            rc = detail::inflate_block(s, t.lit, t.dist);
            if (rc) return rc;
        } else {
            return ERR_CORRUPT;
        }
        if (final_block) break;
    }
    if (consumed) *consumed = (size_t)(s.in - in) - s.bitcnt / 8;   // whole bytes not yet used
    return s.out - out;
}

// zlib stream (2-byte header, deflate data, Adler-32) -> out; bytes written or an error
inline int64_t zlib_decompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, Tables &t)
{
    if (n < 6) return ERR_TRUNCATED;
    const unsigned cmf = in[0], flg = in[1];
    if ((cmf & 0x0F) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20))
        return ERR_HEADER;
    size_t used = 0;
    const int64_t got = inflate_raw(in + 2, n - 2, out, cap, t, &used);
    if (got < 0) return got;
    if (n - 2 - used < 4) return ERR_TRUNCATED;
    const uint8_t *a = in + 2 + used;
    const uint32_t want = ((uint32_t)a[0] << 24) | ((uint32_t)a[1] << 16) | ((uint32_t)a[2] << 8) | a[3];
    if (detail::adler32(out, (size_t)got) != want) return ERR_CHECKSUM;
    return got;
}

}  // namespace pbinf
