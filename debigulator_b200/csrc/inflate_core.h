// inflate_core.h -- warp-per-stream DEFLATE decoder (device code, sm_100a).
//
// Replaces the whole of the reference's inflate() (inflate.c:786-1965): bit
// reader (:225-278, :367-413), stored / fixed / dynamic block parsing
// (:919-989, :1018-1181, :1182-1667), canonical Huffman build (unpack_huffman
// :565-706 + huffman_to_hashmap :494-557), symbol decode (:421-474) and the
// LZ77 copy (:1861-1897). It is a new design, not a translation:
//
//   * one warp owns one stream. Headers are read uniformly (every lane holds the
//     same window registers); symbols are decoded a 32-bit window at a time:
//     lane k decodes the candidate symbol that would start at bit k (litlen
//     lookup, extra bits, distance lookup -- 32 offsets in parallel) and the
//     warp then walks the chain of real symbol starts with one shuffle per
//     symbol. The lanes also fan out for Huffman table construction
//     (match_any ranking, warp scans), LZ77 copies (one byte per lane, stores
//     deferred behind the next symbols, pattern replication for short
//     distances) and stored-block copies.
//   * compressed bytes are staged global -> shared with 16-byte cp.async
//     (LDGSTS) into a 2 x 512 B per-warp ring, requested one chunk ahead.
//   * decode tables are two-level: a 9-bit (litlen) / 8-bit (distance) primary
//     LUT with pre-baked base/extra-bit fields, and a canonical first-code
//     walk for the rare longer codes -- 5.2 KB of shared memory per warp
//     instead of the reference's 3 x 792 KB hash maps (inflate.c:112-118).
//   * when symbols only have to be recorded (the sizes pass of the block-split
//     path) a Huffman block is decoded lane-parallel instead: one lane per
//     sub-chunk behind exact merge points (lane_round below). Streams
//     that are one fixed-Huffman block are decoded by fx_core.h, one lane per
//     chunk.
//
// Behavioural parity with the reference's silent, no-assert build is kept where
// the reference has defined behaviour: the premature
// end-of-stream rule Q2 (inflate.c:1702-1717), the "code length >= alphabet
// size" rejection Q3 (:599-602), the min/max code-length bookkeeping Q5
// (:528-539), BTYPE 3 skipped as an empty block (:990-998), repeat codes
// running across the litlen/dist boundary (:1412-1416), and the argument checks
// (:826-844). Where the reference has undefined behaviour (output overflow,
// reads past the input, symbols 286/287, distance symbols 30/31 handled at
// :1809) this decoder fails the stream instead.
#pragma once
#include "simt.h"

namespace dbg {

// Per-stream status; `good` in the reference's sense is (status == ST_OK).
enum InflateStatus : uint32_t {
    ST_OK = 0,
    ST_CAP_LT_INPUT = 1,    // inflate.c:826 recipient_size < compressed_input_size
    ST_INPUT_TOO_SMALL = 2, // inflate.c:836 compressed_input_size < 5
    ST_TOO_LARGE = 3,       // beyond this implementation's 31/32-bit stream limits
    ST_STORED_LEN = 4,      // inflate.c:949 LEN != ~NLEN
    ST_BAD_TABLE = 5,       // unpack_huffman failure (:599) or over-subscribed code
    ST_BAD_CODE = 6,        // hashed_huffman_decode failure (:465-473)
    ST_BAD_SYMBOL = 7,      // litlen 286/287, distance > 29 (:1809)
    ST_BAD_DISTANCE = 8,    // inflate.c:1843 distance reaches before the output start
    ST_OUT_OVERFLOW = 9,    // output capacity exceeded (UB in the reference)
    ST_TRUNCATED = 10,      // input exhausted at a block boundary / inside a stored block
    ST_BAD_REPEAT = 11,     // code-length repeat with no previous length
    ST_CONTAINER = 12,      // set by container parsers (PNG / gzip), not by inflate
};

struct InflateSmem {
    uint32_t ring[256];      // 2 x 512 B input ring
    uint32_t lit_lut[512];   // 9-bit primary litlen table
    uint32_t dist_lut[256];  // 8-bit primary distance table (its first 128 entries double as the code-length code table)
    uint16_t lit_sorted[288];
    uint16_t lit_first[16], lit_offs[16], lit_cnt[16];
    uint16_t dist_first[16], dist_offs[16], dist_cnt[16];
    uint8_t dist_sorted[32];
    uint8_t lens[320];       // HLIT (<=288) + HDIST (<=32) code lengths
    uint32_t cand[2][64];    // candidate symbols of the current / previous decode pass, one per start offset
};

// LUT entry: [3:0] code length (0 = not in the primary table), bit4 literal /
// plain value, bit5 end-of-block, bit6 length-or-distance with base in [31:16]
// and extra-bit count in [12:8], bit7 undecodable symbol.
enum { E_LIT = 0x10, E_EOB = 0x20, E_BASE = 0x40, E_BAD = 0x80 };
enum { K_CLEN = 0, K_LITLEN = 1, K_DIST = 2 };
enum { LIT_ROOT = 9, DIST_ROOT = 8 };

template <int KIND>
DBG_DEV uint32_t make_entry(uint32_t sym, uint32_t l)
{
    if (KIND == K_CLEN) return (sym << 16) | E_LIT | l;
    if (KIND == K_LITLEN) {
        if (sym < 256) return (sym << 16) | E_LIT | l;
        if (sym == 256) return E_EOB | l;
        if (sym > 285) return E_BAD | l;
        uint32_t i = sym - 257, xb, base;  // inflate.c:716-746
        if (i < 8) { xb = 0; base = 3 + i; }
        else if (i == 28) { xb = 0; base = 258; }
        else { xb = (i >> 2) - 1; base = 3 + ((4 + (i & 3)) << xb); }
        return (base << 16) | (xb << 8) | E_BASE | l;
    }
    if (sym > 29) return E_BAD | l;        // inflate.c:1809
    uint32_t xb = sym < 4 ? 0 : (sym >> 1) - 1;  // inflate.c:748-779
    uint32_t base = sym < 4 ? sym + 1 : 1 + ((2 + (sym & 1)) << xb);
    return (base << 16) | (xb << 8) | E_BASE | l;
}

// ---------------------------------------------------------------- bit reader --
// A 128-bit window (4 words, identical in every lane) over the shared-memory
// input ring. `s` is the bit offset of the next unread bit inside w0. The same
// structure serves the serial header parsing (peek32 / consume, uniform) and
// the lane-parallel symbol decode, where lane k looks at the stream from bit
// offset k of the window.
struct Window {
    const uint8_t *base;  // 16-byte aligned global address at or below the stream start
    uint32_t *ring;
    uint32_t end16;       // ring-coordinate byte offset past which input reads as zero
    uint32_t w0, w1, w2, w3;
    uint32_t wb;          // ring-coordinate index of the word held in w0
    uint32_t s;           // 0..31
    int32_t wleft;        // words until the one holding the limit bit (<= 0: the limit is in or before w0)
    uint32_t lim_w, lim_b;  // ring-coordinate word index / bit of the limit: no symbol may start at or past it

    DBG_DEVM void load_chunk(uint32_t c)
    {
        uint32_t off = (c << 9) + ((uint32_t)simt::lane() << 4);
        bool in = off < end16;
        simt::cp_async16_stream(&ring[((c & 1) << 7) + ((uint32_t)simt::lane() << 2)], base + (in ? off : 0), in ? 16 : 0);
    }
    DBG_DEVM void maintain()
    {
        if (wb & 64) {  // half way through a chunk: the next one has landed
            simt::cp_async_wait_all();
            simt::syncwarp();
        } else {        // entered a new chunk: recycle the slot behind us
            simt::syncwarp();
            load_chunk((wb >> 7) + 1);
            simt::cp_async_commit();
        }
    }
    DBG_DEVM void shift()
    {
        w0 = w1;
        w1 = w2;
        w2 = w3;
        wb++;
        wleft--;
        if ((wb & 63) == 0) maintain();
        w3 = ring[(wb + 3) & 255];
    }
    DBG_DEVM uint32_t peek32() const { return simt::funnel_r(w0, w1, s); }
    DBG_DEVM void consume(uint32_t n)  // n <= 64
    {
        s += n;
        while (s >= 32) {
            s -= 32;
            shift();
        }
    }
    DBG_DEVM uint64_t abs_bits() const { return ((uint64_t)wb << 5) + s; }
    DBG_DEVM void seek(uint64_t bytepos)
    {
        uint32_t c = (uint32_t)(bytepos >> 9);
        simt::cp_async_wait_all();  // nothing from an earlier position may land after this
        simt::syncwarp();
        load_chunk(c);
        load_chunk(c + 1);
        simt::cp_async_commit();
        simt::cp_async_wait_all();
        simt::syncwarp();
        wb = (uint32_t)(bytepos >> 2);
        w0 = ring[wb & 255];
        w1 = ring[(wb + 1) & 255];
        w2 = ring[(wb + 2) & 255];
        w3 = ring[(wb + 3) & 255];
        s = ((uint32_t)bytepos & 3) << 3;
        wleft = (int32_t)(lim_w - wb);
    }
    DBG_DEVM void seek_bits(uint64_t bitpos)
    {
        seek(bitpos >> 3);
        consume((uint32_t)bitpos & 7);
    }
    DBG_DEVM void set_limit(uint64_t limit_bits)
    {
        lim_w = (uint32_t)(limit_bits >> 5);
        lim_b = (uint32_t)limit_bits & 31;
        wleft = (int32_t)(lim_w - wb);
    }
};

// ------------------------------------------------------------ table builder --
// Canonical Huffman construction for `n` code lengths (RFC 1951 3.2.2 as in
// unpack_huffman inflate.c:565-706), executed by the whole warp:
//   pass 1  per-length histogram via match_any, reference validity rule
//           (any length >= n fails, :599-602) and the reference's effective
//           maximum code length (Q5, :528-539) via a warp prefix-min;
//   scan    first code and sorted-order offset of every length (lane l owns
//           length l);
//   pass 2  stable counting sort of the symbols by (length, symbol);
//   pass 3  primary LUT fill from the sorted order (bit-reversed replication).
template <int KIND, typename SORTED_T>
DBG_DEV bool build_table(const uint8_t *lens, int n, int root, uint32_t *lut, SORTED_T *sorted, uint16_t *first,
                         uint16_t *offs, uint16_t *cnt, uint32_t *eff_max_out)
{
    const int ln = simt::lane();
    if (ln < 16) cnt[ln] = 0;
    for (int i = ln; i < (1 << root); i += 32) lut[i] = 0;
    simt::syncwarp();

    uint32_t run_min = 99, emax = 1;
    bool bad = false;
    for (int base = 0; base < n; base += 32) {
        int s = base + ln;
        uint32_t l = (s < n) ? lens[s] : 0;
        bad |= (l >= (uint32_t)n);
        uint32_t m = simt::match_any(l);
        if (l && ln == 31 - simt::clz(m)) cnt[l] = (uint16_t)(cnt[l] + simt::popc(m));
        uint32_t pm = l ? l : 99;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = simt::shfl_up(pm, d);
            if (ln >= d && t < pm) pm = t;
        }
        uint32_t excl = simt::shfl_up(pm, 1);
        if (ln == 0) excl = 99;
        if (run_min < excl) excl = run_min;
        uint32_t cand = (l && excl <= l) ? l : 0;
        for (int d = 16; d; d >>= 1) {
            uint32_t t = simt::shfl_xor(cand, d);
            if (t > cand) cand = t;
        }
        if (cand > emax) emax = cand;
        uint32_t last = simt::shfl(pm, 31);
        if (last < run_min) run_min = last;
        simt::syncwarp();
    }
    if (simt::any(bad)) return false;

    uint32_t c_l = (ln >= 1 && ln < 16) ? cnt[ln] : 0;
    uint32_t incl = c_l;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = simt::shfl_up(incl, d);
        if (ln >= d) incl += t;
    }
    uint32_t code = 0;
    for (int j = 1; j < 16; j++) {
        uint32_t cj = simt::shfl(c_l, j);
        if (j < ln && ln < 16) code += cj << (ln - j);
    }
    bool over = (ln >= 1 && ln < 16 && c_l && code + c_l > (1u << ln));
    if (ln < 16) {
        first[ln] = (uint16_t)code;
        offs[ln] = (uint16_t)(incl - c_l);
    }
    uint32_t total = simt::shfl(incl, 15);
    if (simt::any(over)) return false;
    simt::syncwarp();

    for (int base = 0; base < n; base += 32) {
        int s = base + ln;
        uint32_t l = (s < n) ? lens[s] : 0;
        uint32_t m = simt::match_any(l);
        uint32_t b = l ? offs[l] : 0;
        simt::syncwarp();
        if (l) {
            if (ln == 31 - simt::clz(m)) offs[l] = (uint16_t)(b + simt::popc(m));
            sorted[b + simt::popc(m & ((1u << ln) - 1))] = (SORTED_T)s;
        }
        simt::syncwarp();
    }
    if (ln >= 1 && ln < 16) offs[ln] = (uint16_t)(offs[ln] - cnt[ln]);
    simt::syncwarp();

    for (uint32_t base = 0; base < total; base += 32) {
        uint32_t k = base + ln;
        if (k < total) {
            uint32_t s = sorted[k];
            uint32_t l = lens[s];
            if (l <= (uint32_t)root && l <= emax) {
                uint32_t c = first[l] + (k - offs[l]);
                uint32_t rev = simt::brev(c) >> (32 - l);
                uint32_t e = make_entry<KIND>(s, l);
                for (uint32_t j = rev; j < (1u << root); j += (1u << l)) lut[j] = e;
            }
        }
    }
    simt::syncwarp();
    *eff_max_out = emax;
    return true;
}

// Codes longer than the primary index: canonical walk, shortest length first,
// which is the order the reference probes in (inflate.c:437-463).
template <int KIND, typename SORTED_T>
DBG_DEV_NOINLINE uint32_t slow_decode(uint32_t bits, int root, uint32_t maxlen, const SORTED_T *sorted,
                                      const uint16_t *first, const uint16_t *offs, const uint16_t *cnt)
{
    uint32_t v = simt::brev(bits);
    for (uint32_t l = (uint32_t)root + 1; l <= maxlen; l++) {
        uint32_t idx = (v >> (32 - l)) - first[l];
        if (idx < cnt[l]) return make_entry<KIND>(sorted[offs[l] + idx], l);
    }
    return 0;
}

// ------------------------------------------------------------------- copies --
// One deferred store per lane: the first <=32 bytes of a match are loaded when
// the match is decoded and written when the NEXT match (or the end of the
// stream) needs them, so the global-load latency overlaps the decode of the
// following symbols instead of stalling the warp.
struct PendingStore {
    uint8_t *ptr;   // byte sink: destination byte; 16-bit sink: destination cell
    uint32_t val;
    bool on;
};

DBG_DEV void flush_pending(PendingStore &pd)
{
    if (pd.on) *pd.ptr = (uint8_t)pd.val;
    pd.on = false;
}
DBG_DEV void flush_pending16(PendingStore &pd)
{
    if (pd.on) *reinterpret_cast<uint16_t *>(pd.ptr) = (uint16_t)pd.val;
    pd.on = false;
}

// LZ77 match, general case (inflate.c:1861-1897): longer than one 32-byte
// chunk and / or overlapping. The caller guarantees dist <= pos and
// pos + len <= cap and has flushed the pending store. All lanes participate.
DBG_DEV_NOINLINE void copy_match_slow(uint8_t *out, uint32_t pos, uint32_t len, uint32_t dist)
{
    const uint32_t ln = (uint32_t)simt::lane();
    uint8_t *dst = out + pos;
    const uint8_t *src = dst - dist;
    simt::syncwarp();  // earlier stores by other lanes are visible from here on
    if (dist >= len) {
        // no overlap: all loads first, then all stores (a store-load-store chain would pay one L2 round
        // trip per 32 bytes; len <= 258 means at most 9 bytes per lane)
        uint32_t v[9];
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const uint32_t i = ln + 32 * j;
            v[j] = i < len ? src[i] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const uint32_t i = ln + 32 * j;
            if (i < len) dst[i] = (uint8_t)v[j];
        }
    } else if (dist >= 32) {
        for (uint32_t b = 0; b < len; b += 32) {
            uint32_t i = b + ln;
            if (i < len) dst[i] = src[i];
            simt::syncwarp();
        }
    } else if (len >= 48 && (dist == 1 || dist == 2 || dist == 4)) {
        // run-length train (a run of one byte, one 16-bit sample or one RGBA pixel: what flat image areas and zero runs
        // deflate to). The period divides 4, so every aligned output word holds the same rotation of the pattern: head
        // bytes up to a 16-byte boundary, then one 16-byte store per lane and 512 bytes -- a 258-byte match is one store
        // instruction instead of nine shuffle + byte-store rounds (gimp_test.png is 20 K such matches).
        uint32_t p4;
        if (dist == 1) p4 = (uint32_t)src[0] * 0x01010101u;
        else if (dist == 2) p4 = ((uint32_t)src[0] | ((uint32_t)src[1] << 8)) * 0x00010001u;
        else p4 = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
        const uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);  // < len
        if (ln < head) dst[ln] = (uint8_t)(p4 >> (8 * (ln & 3)));
        const uint32_t pat = simt::funnel_r(p4, p4, 8 * (head & 3));
        const uint32_t body = (len - head) & ~15u;
        for (uint32_t o = 16 * ln; o < body; o += 512) simt::st_u32x4((uint32_t *)(dst + head + o), pat, pat, pat, pat);
        const uint32_t done = head + body;
        if (ln < len - done) dst[done + ln] = (uint8_t)(p4 >> (8 * ((done + ln) & 3)));
    } else {
        // overlapping short distance: replicate the dist-byte pattern from registers
        uint32_t idx = ln % dist;
        uint32_t v = src[idx];
        uint32_t step = 32 % dist;
        for (uint32_t b = 0; b < len; b += 32) {
            uint32_t x = simt::shfl(v, (int)idx);
            if (b + ln < len) dst[b + ln] = (uint8_t)x;
            idx += step;
            if (idx >= dist) idx -= dist;
        }
    }
}

// Stored block payload (inflate.c:958-989): a plain copy with arbitrary source and destination alignment.
// Head bytes bring the destination to a 4-byte boundary, then every lane moves one word per step, built
// from two aligned source words with a funnel shift.
DBG_DEV void copy_stored(uint8_t *dst, const uint8_t *src, uint32_t len)
{
    const uint32_t ln = (uint32_t)simt::lane();
    uint32_t head = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    if (head > len) head = len;
    if (ln < head) dst[ln] = src[ln];
    dst += head;
    src += head;
    len -= head;
    const uint32_t words = len >> 2;
    const uint32_t sh = (uint32_t)((uintptr_t)src & 3) * 8;
    const uint32_t *s4 = (const uint32_t *)((uintptr_t)src & ~(uintptr_t)3);
    uint32_t *d4 = (uint32_t *)dst;
    // four words per lane and step, all loads before the stores: one memory round trip per 512 bytes instead of per 128
    // (ncu: 57 % of the kernel's stall samples sat on the funnel shift of the one-word loop when stored members dominate)
    uint32_t i = ln;
    for (; i + 96 < words; i += 128) {
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            lo[j] = s4[i + 32 * j];
            hi[j] = sh ? s4[i + 32 * j + 1] : 0u;  // s4[.. + 1] still overlaps the source when sh != 0
        }
#pragma unroll
        for (int j = 0; j < 4; j++) d4[i + 32 * j] = sh ? simt::funnel_r(lo[j], hi[j], sh) : lo[j];
    }
    for (; i < words; i += 32) {
        uint32_t lo = s4[i];
        uint32_t v = sh ? simt::funnel_r(lo, s4[i + 1], sh) : lo;
        d4[i] = v;
    }
    const uint32_t done = words << 2;
    if (ln < len - done) dst[done + ln] = src[done + ln];
}

// Match dispatch: the common short non-overlapping match becomes a deferred
// load/store pair; everything else goes through copy_match_slow.
DBG_DEV void copy_match(uint8_t *out, uint32_t pos, uint32_t len, uint32_t dist, PendingStore &pd)
{
    flush_pending(pd);
    if ((len <= 32) & (dist >= len)) {
        simt::syncwarp();  // earlier stores by other lanes are visible from here on
        const uint8_t *sp = out + (pos - dist + (uint32_t)simt::lane());
        pd.on = (uint32_t)simt::lane() < len;
        if (pd.on) pd.val = *sp;
        pd.ptr = const_cast<uint8_t *>(sp) + dist;
    } else {
        copy_match_slow(out, pos, len, dist);
    }
}

DBG_DEV uint32_t swizzle_at(uint32_t i)
{
    // code-length alphabet order 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15 (inflate.c:25-26)
    const uint64_t lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) |
                        (9ull << 30) | (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
    const uint64_t hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
    return (uint32_t)((i < 12 ? lo >> (5 * i) : hi >> (5 * (i - 12))) & 31);
}

// --------------------------------------------------- lane-parallel symbol decode --
// Candidate word of lane k = "the symbol that would start at bit k of the
// window": [6:0] window offset of the next symbol (<= 79), [15:7] match length
// (0 = literal, 1 = special), [31:16] literal byte | distance-1 | special code.
enum { CAND_EOB = 0, CAND_ERR = 1, CAND_SLOW = 2 };
DBG_DEV uint32_t cand_special(uint32_t next, uint32_t code) { return next | (1u << 7) | (code << 16); }

DBG_DEV uint32_t decode_candidate(const InflateSmem *sm, uint32_t lo, uint32_t mid, uint32_t k)
{
    uint32_t e = sm->lit_lut[lo & ((1u << LIT_ROOT) - 1)];
    uint32_t l1 = e & 15;
    if (e & E_LIT) return (k + l1) | (e & 0xffff0000u);
    if (e & E_BASE) {
        uint32_t xb = (e >> 8) & 31;
        uint32_t len = (e >> 16) + ((lo >> l1) & ((1u << xb) - 1));
        uint32_t t1 = l1 + xb;
        uint32_t v = simt::funnel_r(lo, mid, t1);
        uint32_t e2 = sm->dist_lut[v & ((1u << DIST_ROOT) - 1)];
        uint32_t l2 = e2 & 15;
        if (e2 & E_BASE) {
            uint32_t xb2 = (e2 >> 8) & 31;
            uint32_t dist = (e2 >> 16) + ((v >> l2) & ((1u << xb2) - 1));
            return (k + t1 + l2 + xb2) | (len << 7) | ((dist - 1) << 16);
        }
        return cand_special(32, l2 == 0 ? CAND_SLOW : CAND_ERR);
    }
    if (l1 == 0) return cand_special(32, CAND_SLOW);
    if (e & E_EOB) return cand_special(k + l1, CAND_EOB);
    return cand_special(32, CAND_ERR);
}

// ------------------------------------------------------------- output sinks --
// The symbol decoder is shared by three consumers:
//   SINK_BYTES  the normal decode: bytes to the output buffer (deferred match stores);
//   SINK_COUNT  chunk probing for the split-stream path: only sizes are counted;
//   SINK_U16    chunk decode for the split-stream path: 16-bit cells, where a value
//               >= 256 is a marker "byte at distance 32768 - (v - 256) before this chunk's
//               output start", resolved later against the finished output.
//   SINK_TOKENS the sizes-only pass of the block-split path, which also records every symbol as a
//               32-bit token (literal: the byte; match: bit 31 | length << 16 | distance - 1) so that
//               the second pass expands tokens instead of decoding Huffman codes again.
enum { SINK_BYTES = 0, SINK_COUNT = 1, SINK_U16 = 2, SINK_TOKENS = 3 };
constexpr uint32_t TOKEN_MATCH = 0x80000000u;
enum { END_EOB = 0, END_LIMIT = 1 };

struct Sink {
    uint8_t *out;     // SINK_BYTES
    uint16_t *out16;  // SINK_U16
    uint32_t pos;     // bytes produced so far (relative to out / out16)
    uint32_t cap;
    uint64_t abs_base;  // SINK_U16: stream output offset of out16[0] (markers may not reach before the stream)
    uint32_t *tok;      // SINK_TOKENS: token area, `tok_cap` entries; ntok keeps counting past it (= overflow)
    uint32_t ntok, tok_cap;
    bool lanes;         // SINK_TOKENS / SINK_BYTES: Huffman blocks may be decoded lane-parallel (lane_round)
    uint32_t round_bits;  // SINK_BYTES: round length
    uint32_t lane_bad, lane_skip;  // back-off: failures in a row / blocks still to leave to the symbol walk
    uint32_t *lb_stats; // optional counters of lane_round: [0] rounds, [1] rounds that met end-of-block, [2] rounds that did not
    PendingStore pd;
};

// CHECK = false: the caller has verified that the output has room for everything one window can produce.
template <int SINK, bool CHECK>
DBG_DEV uint32_t emit_literal(Sink &k, uint32_t byte)
{
    if (SINK == SINK_TOKENS) {
        if (k.ntok < k.tok_cap && simt::lane() == 0) k.tok[k.ntok] = byte;
        k.ntok++;
    } else if (SINK != SINK_COUNT) {
        if (CHECK && k.pos >= k.cap) return ST_OUT_OVERFLOW;
        if (simt::lane() == 0) {
            if (SINK == SINK_BYTES) k.out[k.pos] = (uint8_t)byte;
            else k.out16[k.pos] = (uint16_t)byte;
        }
    }
    k.pos++;
    return ST_OK;
}

DBG_DEV_NOINLINE void copy_match_u16(uint16_t *o, uint32_t pos, uint32_t len, uint32_t dist)
{
    const uint32_t ln = (uint32_t)simt::lane();
    simt::syncwarp();
    // every source index lies before `pos` (overlaps replicate the dist-long pattern), so all
    // reads precede all writes of this match; negative indices are markers into the 32 KiB
    // window that ends where this chunk's output starts
    uint32_t idx = dist >= len ? ln : ln % dist;
    const uint32_t step = dist >= len ? 32 : 32 % dist;
    for (uint32_t b = 0; b < len; b += 32) {
        if (b + ln < len) {
            int32_t si = (int32_t)pos - (int32_t)dist + (int32_t)idx;
            uint32_t v = si < 0 ? (uint32_t)(256 + 32768 + si) : o[si];
            o[pos + b + ln] = (uint16_t)v;
        }
        idx += step;
        if (dist < len && idx >= dist) idx -= dist;
    }
    simt::syncwarp();
}

template <int SINK, bool CHECK>
DBG_DEV uint32_t emit_match(Sink &k, uint32_t len, uint32_t dist)
{
    if (SINK == SINK_TOKENS) {
        if (k.ntok < k.tok_cap && simt::lane() == 0) k.tok[k.ntok] = TOKEN_MATCH | (len << 16) | (dist - 1);
        k.ntok++;
    }
    if (SINK == SINK_COUNT || SINK == SINK_TOKENS) {
        k.pos += len;
        return ST_OK;
    }
    if (CHECK && k.pos + len > k.cap) return ST_OUT_OVERFLOW;
    if (SINK == SINK_BYTES) {
        if (dist > k.pos) return ST_BAD_DISTANCE;  // inflate.c:1843
        copy_match(k.out, k.pos, len, dist, k.pd);
    } else {
        if (dist > 32768 || dist > k.abs_base + k.pos) return ST_BAD_DISTANCE;
        flush_pending16(k.pd);
        if ((len <= 32) & (dist >= len)) {
            // same deferred load/store pair as the byte sink; a source before the chunk's own output is a marker
            simt::syncwarp();
            const uint32_t ln = (uint32_t)simt::lane();
            const int32_t si = (int32_t)k.pos - (int32_t)dist + (int32_t)ln;
            k.pd.on = ln < len;
            if (k.pd.on) k.pd.val = si < 0 ? (uint32_t)(256 + 32768 + si) : k.out16[si];
            k.pd.ptr = reinterpret_cast<uint8_t *>(k.out16 + k.pos + ln);
        } else {
            copy_match_u16(k.out16, k.pos, len, dist);
        }
    }
    k.pos += len;
    return ST_OK;
}

// ------------------------------------------------------------- block headers --
struct BlockTables {
    uint32_t lit_max, dist_max;
};

// Builds the decode tables of a fixed (btype 1) or dynamic (btype 2) block whose
// 3 header bits have just been consumed (inflate.c:1018-1181 / :1182-1667).
DBG_DEV uint32_t read_huffman_tables(Window &w, InflateSmem *sm, uint32_t btype, BlockTables &bt)
{
    const uint32_t ln = (uint32_t)simt::lane();
    uint32_t hlit = 288, hdist = 32;
    if (btype == 1) {
        // fixed code (inflate.c:1035-1084); distances are 5-bit codes (:1783-1788)
        for (uint32_t i = ln; i < 320; i += 32)
            sm->lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5);
        simt::syncwarp();
    } else {
        uint32_t v = w.peek32();
        hlit = (v & 31) + 257;
        hdist = ((v >> 5) & 31) + 1;
        uint32_t hclen = ((v >> 10) & 15) + 4;
        w.consume(14);
        if (ln < 19) sm->lens[ln] = 0;
        simt::syncwarp();
        for (uint32_t i = 0; i < hclen; i++) {
            if (ln == 0) sm->lens[swizzle_at(i)] = (uint8_t)(w.peek32() & 7);
            w.consume(3);
        }
        simt::syncwarp();
        uint32_t cl_max;
        if (!build_table<K_CLEN, uint8_t>(sm->lens, 19, 7, sm->dist_lut, sm->dist_sorted, sm->dist_first, sm->dist_offs,
                                          sm->dist_cnt, &cl_max))
            return ST_BAD_TABLE;
        const uint32_t n = hlit + hdist;
        uint32_t i = 0, prev = 0;
        while (i < n) {
            uint32_t bits = w.peek32();
            uint32_t e = sm->dist_lut[bits & 127];
            uint32_t l = e & 15;
            if (l == 0) return ST_BAD_CODE;
            uint32_t sym = e >> 16;
            if (sym < 16) {
                w.consume(l);
                if (ln == 0) sm->lens[i] = (uint8_t)sym;
                prev = sym;
                i++;
                continue;
            }
            uint32_t rep, val;
            if (sym == 16) {  // inflate.c:1439-1469
                if (i == 0) return ST_BAD_REPEAT;
                rep = 3 + ((bits >> l) & 3);
                w.consume(l + 2);
                val = prev;
            } else if (sym == 17) {  // :1471-1488
                rep = 3 + ((bits >> l) & 7);
                w.consume(l + 3);
                val = 0;
            } else {  // :1490-1509
                rep = 11 + ((bits >> l) & 127);
                w.consume(l + 7);
                val = 0;
            }
            for (uint32_t k = ln; k < rep; k += 32)
                if (i + k < n) sm->lens[i + k] = (uint8_t)val;
            prev = val;
            i += rep;
        }
        simt::syncwarp();
    }
    if (!build_table<K_LITLEN, uint16_t>(sm->lens, (int)hlit, LIT_ROOT, sm->lit_lut, sm->lit_sorted, sm->lit_first,
                                         sm->lit_offs, sm->lit_cnt, &bt.lit_max))
        return ST_BAD_TABLE;
    if (!build_table<K_DIST, uint8_t>(sm->lens + hlit, (int)hdist, DIST_ROOT, sm->dist_lut, sm->dist_sorted, sm->dist_first,
                                      sm->dist_offs, sm->dist_cnt, &bt.dist_max))
        return ST_BAD_TABLE;
    return ST_OK;
}

// ------------------------------------------------------------ symbol decoder --
// Decodes symbols from the window position until end-of-block (END_EOB) or
// until the next symbol would start at or past the window's limit (END_LIMIT).
// Each pass looks at one 32-bit window: every lane decodes the candidate
// symbol at its own bit offset (LUT lookups, extra bits, distance) in
// parallel, then the warp walks the chain of real symbol starts with one
// shuffle per symbol.
template <int SINK>
DBG_DEV uint32_t decode_symbols(Window &w, InflateSmem *sm, const BlockTables &bt, Sink &k, uint32_t &end_reason)
{
    const uint32_t ln = (uint32_t)simt::lane();
    uint32_t flip = 0;
    // A call that follows another one directly (a stretch between two rounds ends at its limit, the next round finds too
    // little left and hands back to the walk) starts on candidate buffer 0 again: no lane may fill it while a slower lane
    // is still walking the previous call's last pass in it.
    simt::syncwarp();
    for (;;) {
        // Two words (64 start offsets) per pass when the limit is not in sight, otherwise one word
        // with the limit applied. Candidate k of the second word carries offsets relative to the first.
        const bool two = w.wleft >= 2;
        uint32_t lim = two ? 64u : 32u;
        if (w.wleft <= 0) {
            lim = w.wleft == 0 ? w.lim_b : 0u;
            if (w.s >= lim) {
                end_reason = END_LIMIT;
                return ST_OK;
            }
        }
        const uint32_t lo = simt::funnel_r(w.w0, w.w1, ln);
        const uint32_t mid = simt::funnel_r(w.w1, w.w2, ln);
        const uint32_t hi = simt::funnel_r(w.w2, w.w3, ln);
        // candidates go through shared memory: the chain walk below reads them with a uniform
        // (broadcast) address, which needs no warp-convergence point per symbol the way a shuffle does
        // (double buffered: a lane that is already in the next pass must not overwrite what a slower
        // lane is still walking; the barrier below keeps lanes at most one pass apart)
        uint32_t *cand = sm->cand[flip];
        flip ^= 1;
        cand[ln] = decode_candidate(sm, lo, mid, ln);
        if (two) cand[ln + 32] = decode_candidate(sm, mid, hi, ln + 32);
        simt::syncwarp();
        uint32_t p = w.s, cur, err = 0;
        bool eob = false, slow = false;
        // one pass yields at most 64 symbols of at most 258 bytes: with that much room left the
        // per-symbol capacity checks are dropped
        const bool roomy = SINK == SINK_COUNT || SINK == SINK_TOKENS || k.cap - k.pos >= 64 * 258;
        // The walk is unrolled four symbols deep: ptxas makes the warp wait for the deferred match load at the
        // first branch after a loop back-edge (not at the store that needs the bytes), so a one-symbol loop would
        // expose that load's latency on every symbol; inside the unrolled body the wait sits on the consumer.
#define DBG_SYM(CHECK)                                                                         \
            cur = p;                                                                           \
            {                                                                                  \
                const uint32_t info = cand[p];                                                 \
                const uint32_t lf = (info >> 7) & 511;                                         \
                p = info & 127;                                                                \
                if (lf == 0) {                                                                 \
                    err = emit_literal<SINK, CHECK>(k, info >> 16);                            \
                    if (CHECK && err) break;                                                   \
                } else if (lf >= 3) {                                                          \
                    err = emit_match<SINK, CHECK>(k, lf, (info >> 16) + 1);                    \
                    if (err) break;                                                            \
                } else {                                                                       \
                    const uint32_t code = info >> 16;                                          \
                    if (code == CAND_EOB) eob = true;                                          \
                    else if (code == CAND_ERR) err = ST_BAD_SYMBOL;                            \
                    else slow = true; /* a code longer than the primary LUT index starts at cur */ \
                    break;                                                                     \
                }                                                                              \
            }
#define DBG_WALK(CHECK)                                                                        \
        do {                                                                                   \
            DBG_SYM(CHECK)                                                                     \
            if (p >= lim) break;                                                               \
            DBG_SYM(CHECK)                                                                     \
            if (p >= lim) break;                                                               \
            DBG_SYM(CHECK)                                                                     \
            if (p >= lim) break;                                                               \
            DBG_SYM(CHECK)                                                                     \
        } while (p < lim)
        if (roomy) DBG_WALK(false);
        else DBG_WALK(true);
#undef DBG_WALK
#undef DBG_SYM
        if (err) return err;
        if (slow) {
            // rare: decode this one symbol serially (uniform), then rebuild the window candidates
            if (cur >= 32) {
                w.shift();
                cur -= 32;
            }
            w.s = cur;
            uint32_t bits = w.peek32();
            uint32_t e = sm->lit_lut[bits & ((1u << LIT_ROOT) - 1)];
            if ((e & 15) == 0) {
                e = slow_decode<K_LITLEN, uint16_t>(bits, LIT_ROOT, bt.lit_max, sm->lit_sorted, sm->lit_first, sm->lit_offs,
                                                    sm->lit_cnt);
                if (!e) return ST_BAD_CODE;
            }
            w.consume(e & 15);
            if (e & E_LIT) {
                err = emit_literal<SINK, true>(k, e >> 16);
                if (err) return err;
                continue;
            }
            if (e & E_EOB) {
                end_reason = END_EOB;
                return ST_OK;
            }
            if (e & E_BAD) return ST_BAD_SYMBOL;
            bits = w.peek32();
            uint32_t xb = (e >> 8) & 31;
            uint32_t len = (e >> 16) + (bits & ((1u << xb) - 1));
            w.consume(xb);
            bits = w.peek32();
            e = sm->dist_lut[bits & ((1u << DIST_ROOT) - 1)];
            if ((e & 15) == 0) {
                e = slow_decode<K_DIST, uint8_t>(bits, DIST_ROOT, bt.dist_max, sm->dist_sorted, sm->dist_first, sm->dist_offs,
                                                 sm->dist_cnt);
                if (!e) return ST_BAD_CODE;
            }
            if (e & E_BAD) return ST_BAD_SYMBOL;
            xb = (e >> 8) & 31;
            uint32_t l2 = e & 15;
            uint32_t dist = (e >> 16) + ((bits >> l2) & ((1u << xb) - 1));
            w.consume(l2 + xb);
            err = emit_match<SINK, true>(k, len, dist);
            if (err) return err;
            continue;
        }
        // the window now sits exactly on the next symbol start (the chunk prober relies on it)
        w.s = p & 31;
        for (uint32_t n = p >> 5; n; n--) w.shift();
        if (eob) {
            end_reason = END_EOB;
            return ST_OK;
        }
        if (!two && lim < 32) {  // the walk ran into the limit
            end_reason = END_LIMIT;
            return ST_OK;
        }
    }
}

// ---------------------------------------------------------------- the stream --
struct StreamIn {
    uint32_t mis;        // in & 15
    uint64_t end_byte;   // ring coordinates
    uint64_t q2_limit;   // ring-coordinate bit: rule Q2 (inflate.c:1702-1717)
};

DBG_DEV StreamIn open_stream(Window &w, InflateSmem *sm, const uint8_t *in, uint64_t in_size)
{
    StreamIn g;
    g.mis = (uint32_t)((uintptr_t)in & 15);
    g.end_byte = (uint64_t)g.mis + in_size;
    // Q2: the stream ends, successfully, as soon as the byte cursor ceil(P/8) has
    // reached in_size, i.e. P >= 8*in_size - 7.
    g.q2_limit = 8 * g.end_byte - 7;
    w.base = in - g.mis;
    w.ring = sm->ring;
    w.end16 = (uint32_t)((g.end_byte + 15) & ~15ull);
    w.lim_w = (uint32_t)(g.q2_limit >> 5);
    w.lim_b = (uint32_t)g.q2_limit & 31;
    return g;
}

// Stored block payload into 16-bit cells (split paths).
DBG_DEV void copy_stored_u16(uint16_t *dst, const uint8_t *src, uint32_t len)
{
    for (uint32_t i = (uint32_t)simt::lane(); i < len; i += 32) dst[i] = src[i];
}

// Per-lane bit reader straight over global memory (every lane is somewhere else in the batch, so there is
// nothing to stage cooperatively; consecutive words of one lane hit the same L1 line).
struct LaneBits {
    const uint32_t *a;  // 4-byte aligned address at or below the stream start
    uint32_t boff;      // bit offset of the stream start inside a[0]
    uint32_t last;      // index of the last word before the 16-byte boundary at or after the stream end (readable by contract,
                        // the same bytes the warp-per-stream reader sees); words past it read as zero
    uint32_t widx;      // index of the word held in `nxt`
    uint32_t nxt;       // the next word, already on its way when the buffer runs low (its latency hides behind a symbol)
    uint32_t nb;        // valid bits in buf (>= 32 whenever a symbol is decoded)
    uint64_t buf;       // next stream bits, LSB first

    DBG_DEVM uint32_t word(uint32_t i) const { return i <= last ? simt::ldg_u32(a + i) : 0u; }
    DBG_DEVM void open(const uint8_t *in, uint64_t in_size)
    {
        const uintptr_t p = (uintptr_t)in;
        a = (const uint32_t *)(p & ~(uintptr_t)3);
        boff = 8 * (uint32_t)(p & 3);
        last = (uint32_t)((((p + in_size + 15) & ~(uintptr_t)15) - (p & ~(uintptr_t)3)) >> 2) - 1;
    }
    DBG_DEVM void seek(uint64_t stream_bit)
    {
        const uint64_t abit = stream_bit + boff;
        widx = (uint32_t)(abit >> 5);
        const uint32_t sh = (uint32_t)abit & 31;
        const uint64_t two = (uint64_t)word(widx) | ((uint64_t)word(widx + 1) << 32);
        buf = two >> sh;
        nb = 64 - sh;
        widx += 2;
        nxt = word(widx);
    }
    DBG_DEVM void refill()
    {
        if (nb <= 32) {
            buf |= (uint64_t)nxt << nb;
            widx++;
            nb += 32;
            nxt = word(widx);
        }
    }
    DBG_DEVM void drop(uint32_t n)
    {
        buf >>= n;
        nb -= n;
    }
};

// Token writer of one lane: single stores up to the first 16-byte boundary, then four tokens per store (every
// lane writes somewhere else, so a 4-byte store costs the memory system what a 16-byte one does).
struct TokOut {
    uint32_t *p;
    uint32_t q0, q1, q2;
    uint32_t k;
    DBG_DEVM void open(uint32_t *dst)
    {
        p = dst;
        q0 = q1 = q2 = 0;
        k = 0;
    }
    DBG_DEVM void put(uint32_t t)
    {
        if (k == 3) {
            simt::st_u32x4(p, q0, q1, q2, t);
            p += 4;
            k = 0;
        } else if (k == 0 && ((uintptr_t)p & 15)) {
            *p++ = t;
        } else {
            q0 = q1;
            q1 = q2;
            q2 = t;
            k++;
        }
    }
    DBG_DEVM void close()
    {
        if (k == 3) *p++ = q0;
        if (k >= 2) *p++ = q1;
        if (k >= 1) *p++ = q2;
        k = 0;
    }
};

// How a decode run over part of a stream ended (chunk-parallel paths).
enum : uint32_t { CH_RUN = 0, CH_EOB = 1, CH_Q2 = 2, CH_IDLE = 3, CH_ERR = 16 };  // CH_ERR + status

// ------------------------------------------------ lane-parallel block decode --
// The symbol walk above spends ~64 warp instructions per symbol because 32 lanes cooperate on ONE position of the
// stream. When the symbols only have to be RECORDED (SINK_TOKENS: the sizes pass of the block-split path; a later
// kernel expands the tokens), the lanes can instead each decode a piece of the block on their own, ~70 thread
// instructions per symbol, i.e. ~2-3 warp instructions:
//
//   extent   the bits from the block's first symbol to where the block is presumed to end (the caller's stop bit =
//            the next block-boundary hint, capped) are cut into up to 32 equal sub-chunks, one per lane.
//   merge    lane j > 0 does not know where a symbol starts in its sub-chunk. A symbol is at most 48 bits long
//            (15 + 5 + 15 + 13), so one of the first 48 bit positions is a real symbol start. The lane sweeps the set
//            of positions reachable from ANY of those 48 starts in increasing order (a 64-bit reach mask anchored at the
//            lowest one: decode there, add "position + symbol length", move on), so every position is decoded once
//            however many of the 48 chains run through it. Huffman chains merge quickly; the sweep stops as soon as ONE
//            reachable position is left: every chain that is still alive runs through it -- the lane's entry point.
//            Nothing is guessed: if the chains have not merged half way through the sub-chunk, the lanes are not used.
//   runs     lane j decodes from its entry point to lane j + 1's, recording tokens into its own stretch of the chunk's
//            token area. It must arrive there EXACTLY (its path entered sub-chunk j + 1 on one of the 48 starts).
//   chain    from lane 0 (the real start of the block) along the lanes until the run that meets end-of-block; what
//            the later lanes decoded (bits behind the end of the block) is dropped. The runs' tokens are moved
//            together. Anything irregular on that chain (bad code, the rule-Q2 limit, a run that misses its target, a
//            token stretch that is too small) makes the caller decode the block again the ordinary way, so the status
//            and every byte are those of decode_symbols().
// The concatenated runs are the sequential decode: each starts where the previous one ended and all use one step
// function with the block's own tables (LUTs + slow_decode for codes longer than the LUT index, rule Q5 included).
constexpr uint32_t LB_MIN_SUB = 2048;         // smallest sub-chunk, bits
constexpr uint32_t LB_MAX_EXTENT = 1u << 20;  // largest presumed extent, bits (128 KiB of compressed data)
constexpr uint32_t LB_TOKENS_STRETCH = 4096;     // count pass of the block-split path: bits the walk decodes between two rounds of one block
constexpr uint32_t LB_NOHINT_EXTENT = 3u << 17;  // presumed extent when the caller has no hint, bits (48 KiB)
constexpr uint32_t LB_ROUND_BITS = 86016;     // round length of the byte sink (10.5 KiB of compressed data, 2,688 bits per lane; measured on cfg2 with the final
                                              // kernel: 41.9 ms at 6 KiB, 40.8 at 8, 36.5 at 9, 34.7 at 10, 33.5 at 10.5, 36.7 at 11, 38.6 at 12 -- the warps' token
                                              // scratch in flight competes with the output for the 126 MB L2)
constexpr uint32_t LB_ROUND_TOKENS = 16384;   // its token scratch per warp (a round of the densest sensible code: 4 bits per symbol)
constexpr uint32_t LB_MERGE_BITS = 1024;      // chains that have not merged after this many bits are given up
constexpr uint32_t LB_STARTS = 48;            // candidate entry offsets per sub-chunk = longest possible symbol
enum : uint32_t { LBK_LIT = 0, LBK_MATCH = 1, LBK_EOB = 2, LBK_BAD = 3 };

// One symbol with the block's tables; consumes it. *nbits = its length, *len = output bytes, *tok = its token.
DBG_DEV uint32_t lb_symbol(const InflateSmem *sm, const BlockTables &bt, LaneBits &br, uint32_t *nbits, uint32_t *len, uint32_t *tok)
{
    br.refill();  // >= 33 bits: literal/length code (<= 15) + extra bits (<= 5)
    uint32_t x = (uint32_t)br.buf;
    uint32_t e = sm->lit_lut[x & ((1u << LIT_ROOT) - 1)];
    if ((e & 15) == 0) {
        e = slow_decode<K_LITLEN, uint16_t>(x, LIT_ROOT, bt.lit_max, sm->lit_sorted, sm->lit_first, sm->lit_offs, sm->lit_cnt);
        if (!e) return LBK_BAD;
    }
    const uint32_t l1 = e & 15;
    *len = 1;
    *tok = e >> 16;
    *nbits = l1;
    if (e & E_LIT) {
        br.drop(l1);
        return LBK_LIT;
    }
    if (e & E_EOB) {
        br.drop(l1);
        return LBK_EOB;
    }
    if (!(e & E_BASE)) return LBK_BAD;  // litlen 286 / 287
    const uint32_t xb = (e >> 8) & 31;
    const uint32_t ln = (e >> 16) + ((x >> l1) & ((1u << xb) - 1));
    br.drop(l1 + xb);
    br.refill();  // distance code (<= 15) + extra bits (<= 13)
    x = (uint32_t)br.buf;
    uint32_t e2 = sm->dist_lut[x & ((1u << DIST_ROOT) - 1)];
    if ((e2 & 15) == 0) {
        e2 = slow_decode<K_DIST, uint8_t>(x, DIST_ROOT, bt.dist_max, sm->dist_sorted, sm->dist_first, sm->dist_offs, sm->dist_cnt);
        if (!e2) return LBK_BAD;
    }
    if (!(e2 & E_BASE)) return LBK_BAD;  // distance symbols 30 / 31
    const uint32_t l2 = e2 & 15, xb2 = (e2 >> 8) & 31;
    const uint32_t dist = (e2 >> 16) + ((x >> l2) & ((1u << xb2) - 1));
    br.drop(l2 + xb2);
    *nbits = l1 + xb + l2 + xb2;
    *len = ln;
    *tok = TOKEN_MATCH | (ln << 16) | (dist - 1);
    return LBK_MATCH;
}

// Entry point of the sub-chunk that starts at ring bit `s0`: the one position all chains from its first LB_STARTS
// bit offsets run through, or ~0 when they have not merged within `give_up` bits (or all of them died).
DBG_DEV uint64_t lb_merge_point(const InflateSmem *sm, const BlockTables &bt, const uint8_t *base16, uint64_t end_byte, uint64_t s0,
                                uint32_t give_up, uint64_t q2_limit)
{
    LaneBits br;
    br.open(base16, end_byte);
    br.seek(s0);
    uint64_t reach = (1ull << LB_STARTS) - 1;  // bit i: position anchor + i is reachable
    uint64_t anchor = s0;
    for (uint32_t steps = 0; steps < 2 * LB_STARTS + give_up / 2; steps++) {
        if (reach == 0) return ~0ull;
        const uint32_t lo32 = (uint32_t)reach;
        const uint32_t p = lo32 ? (uint32_t)simt::ffs(lo32) - 1 : 31 + (uint32_t)simt::ffs((uint32_t)(reach >> 32));
        if ((reach >> p) == 1) return anchor + p;  // one position left
        // move the window to the lowest reachable position
        if (p) {
            uint32_t r = p;
            if (r > 32) {
                br.refill();
                br.drop(32);
                r -= 32;
            }
            br.refill();
            br.drop(r);
            anchor += p;
            reach >>= p;
        }
        if (anchor - s0 > give_up) return ~0ull;
        reach &= ~1ull;
        if (anchor >= q2_limit) continue;  // no symbol may start here (rule Q2): this chain ends
        LaneBits t = br;  // decode without consuming
        uint32_t nbits, len, tok;
        const uint32_t kind = lb_symbol(sm, bt, t, &nbits, &len, &tok);
        if (kind == LBK_LIT || kind == LBK_MATCH) reach |= 1ull << nbits;  // (end-of-block and bad codes end a chain)
    }
    return ~0ull;
}

enum : uint32_t { LB_UNUSED = 0, LB_DONE = 1, LB_PARTIAL = 2 };
#ifdef DBG_SIMT_EMU
static uint32_t g_lb_done = 0, g_lb_tried = 0, g_lb_partial = 0;  // emulator only: outcome counts of lane_round
#endif

struct LaneRound {
    uint32_t ntok;     // tokens written to the area, in order, contiguous
    uint32_t out;      // bytes they produce
    uint64_t resume;   // ring bit where the decode goes on (LB_DONE: behind the end-of-block code)
    uint32_t used;     // runs (lanes) that counted
    uint32_t stride;   // COMPACT = false: run j's tokens start at area + j * stride ...
    uint32_t mine;     // ... and this lane's run has `mine` of them (0 when it does not count)
};

// One lane-parallel round over the Huffman block whose tables are built, from ring bit p0 (a symbol start):
//   LB_DONE     up to and including end-of-block;
//   LB_PARTIAL  a prefix of the block: r.resume is the symbol where the caller goes on (another round, or
//               decode_symbols() with the same tables);
//   LB_UNUSED   nothing was decoded.
// `ext` = presumed bits to the end of the block (from a boundary hint, or simply a round length), `area` / `area_cap`
// = where the tokens go. `stats` (optional): [0] rounds, [1] rounds that met end-of-block, [2] rounds that did not.
template <bool COMPACT>
DBG_DEV uint32_t lane_round(const InflateSmem *sm, const BlockTables &bt, const uint8_t *base16, const StreamIn &g, uint64_t p0, uint64_t ext,
                            uint32_t *area, uint32_t area_cap, LaneRound &r, uint32_t *stats)
{
    const uint32_t ln = (uint32_t)simt::lane();
    r.ntok = 0;
    r.out = 0;
    r.resume = p0;
    r.used = 0;
    r.stride = 0;
    r.mine = 0;
    if (ext > LB_MAX_EXTENT) ext = LB_MAX_EXTENT;
    if (ext < 4 * LB_MIN_SUB) return LB_UNUSED;
    uint32_t sub = ((uint32_t)ext + 31) / 32;
    if (sub < LB_MIN_SUB) sub = LB_MIN_SUB;
    sub = (sub + 31) & ~31u;
    const uint32_t L = ((uint32_t)ext + sub - 1) / sub;  // lanes in use, 4..32
    const uint32_t stride = area_cap / L;                // token slots per lane
    if (stride < sub / 16) return LB_UNUSED;
#ifdef DBG_SIMT_EMU
    if (ln == 0) g_lb_tried++;
#else
    if (ln == 0 && stats) atomicAdd(&stats[0], 1u);
#endif
    // merge points. A lane whose chains do not merge (it may be looking at the NEXT block's bits, coded with other
    // tables) has none; the chain below simply ends before it.
    uint64_t entry = p0;
    const uint32_t give_up = sub / 2 < LB_MERGE_BITS ? sub / 2 : LB_MERGE_BITS;
    if (ln > 0 && ln < L) entry = lb_merge_point(sm, bt, base16, g.end_byte, p0 + (uint64_t)ln * sub, give_up, g.q2_limit);
    if (ln >= L) entry = ~0ull;
    const uint32_t e_lo = simt::shfl_down((uint32_t)entry, 1), e_hi = simt::shfl_down((uint32_t)(entry >> 32), 1);
    uint64_t target = ((uint64_t)e_hi << 32) | e_lo;  // the next lane's entry point
    const bool has_target = ln + 1 < L && target != ~0ull;
    // a run without one goes to the first symbol boundary behind its own sub-chunk (it may meet end-of-block on the way)
    if (!has_target) target = p0 + (uint64_t)(ln + 1) * sub;
    // runs
    enum : uint32_t { R_ARRIVED = 0, R_EOB = 1, R_STOPPED = 2, R_BROKEN = 3 };
    uint32_t n = 0, out = 0, flag = R_BROKEN;
    uint64_t pos = entry;
    if (entry != ~0ull) {
        LaneBits br;
        br.open(base16, g.end_byte);
        br.seek(entry);
        TokOut wr;
        wr.open(area + (uint64_t)ln * stride);
        flag = R_ARRIVED;
        while (pos < target) {
            if (pos >= g.q2_limit || n >= stride) {
                flag = R_BROKEN;  // rule Q2 / no room: the ordinary decoder takes over from this run's entry point
                break;
            }
            uint32_t nbits, len, tok;
            const uint32_t kind = lb_symbol(sm, bt, br, &nbits, &len, &tok);
            if (kind == LBK_BAD) {
                flag = R_BROKEN;
                break;
            }
            pos += nbits;
            if (kind == LBK_EOB) {
                flag = R_EOB;
                break;
            }
            wr.put(tok);
            n++;
            out += len;
        }
        wr.close();
        if (flag == R_ARRIVED) {
            if (!has_target) flag = R_STOPPED;           // a clean prefix: stands on a symbol boundary
            else if (pos != target) flag = R_BROKEN;     // missed the next entry point (cannot happen on the real chain)
        }
    }
    // chain: lanes 0 .. the first one that did not simply arrive at the next entry point
    const uint32_t stops = simt::ballot(flag != R_ARRIVED);  // never empty: the last lane has no target
    const uint32_t stop_lane = (uint32_t)simt::ffs(stops) - 1;
    const uint32_t stop_flag = simt::shfl(flag, (int)stop_lane);
    const uint32_t used = stop_flag == R_BROKEN ? stop_lane : stop_lane + 1;  // runs that count
    if (used == 0) return LB_UNUSED;
    r.used = used;
    r.stride = stride;
    uint32_t in = ln < used ? n : 0u, io = ln < used ? out : 0u;
    const uint32_t mine = in;
    r.mine = mine;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t yn = simt::shfl_up(in, d), yo = simt::shfl_up(io, d);
        if (ln >= (uint32_t)d) {
            in += yn;
            io += yo;
        }
    }
    r.ntok = simt::shfl(in, 31);
    r.out = simt::shfl(io, 31);
    // where the decode goes on: behind the last run that counts, or on the entry point of the broken one
    const uint64_t res = stop_flag == R_BROKEN ? entry : pos;
    r.resume = ((uint64_t)simt::shfl((uint32_t)(res >> 32), (int)stop_lane) << 32) | simt::shfl((uint32_t)res, (int)stop_lane);
    // move the runs' tokens together (run 0 is in place); forward copies, the loads of a step before its stores
    simt::syncwarp();
    for (uint32_t j = 1; COMPACT && j < used; j++) {
        const uint32_t cnt = simt::shfl(mine, (int)j), dst0 = simt::shfl(in - mine, (int)j);
        const uint32_t *src = area + (uint64_t)j * stride;
        // 128 tokens per step, the four loads of a lane before its stores: a run moves DOWN (dst0 + i < j * stride + i), so a
        // store can only land on a source that was read in this step or an earlier one, never on a later one -- and one
        // memory round trip now moves 128 tokens instead of 32 (the loop is a pure latency chain: tokens written a moment
        // ago by other lanes come back through L2)
        for (uint32_t i0 = 0; i0 < cnt; i0 += 128) {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = i0 + 32 * k + ln;
                v[k] = i < cnt ? src[i] : 0u;
            }
            simt::syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = i0 + 32 * k + ln;
                if (i < cnt) area[dst0 + i] = v[k];
            }
            simt::syncwarp();
        }
    }
    const bool whole = stop_flag == R_EOB;
#ifdef DBG_SIMT_EMU
    if (ln == 0) (whole ? g_lb_done : g_lb_partial)++;
#else
    if (ln == 0 && stats) atomicAdd(&stats[whole ? 1 : 2], 1u);
#endif
    return whole ? LB_DONE : LB_PARTIAL;
}

// Tokens -> bytes, by the warp that decodes the stream in order (so every source byte is final): 32 tokens per step,
// an exclusive scan of their lengths gives each its place, literals are stored at once, matches whose source lies wholly
// before the step's output (and that are short) are copied by their own lanes side by side, the others one after the
// other by the whole warp. Returns ST_OK or the status of the FIRST token (in stream order) that fails: a distance that
// reaches before the output start (inflate.c:1843) or an overflow of the capacity.
DBG_DEV uint32_t expand_tokens_bytes(const uint32_t *tok, uint32_t ntok, uint8_t *out, uint32_t &pos_io, uint32_t cap)
{
    const uint32_t ln = (uint32_t)simt::lane();
    uint32_t pos = pos_io;
    uint32_t t_next = ln < ntok ? tok[ln] : 0u;
    for (uint32_t base = 0; base < ntok; base += 32) {
        const bool have = base + ln < ntok;
        const uint32_t t = t_next;
        t_next = base + 32 + ln < ntok ? tok[base + 32 + ln] : 0u;  // the next step's tokens are on their way
        const bool is_match = have && (t & TOKEN_MATCH);
        const uint32_t len = !have ? 0u : is_match ? (t >> 16) & 0x1ff : 1u;
        const uint32_t dist = (t & 0x7fff) + 1;
        uint32_t incl = len;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = simt::shfl_up(incl, d);
            if (ln >= (uint32_t)d) incl += y;
        }
        const uint32_t o = pos + incl - len;
        const uint32_t total = simt::shfl(incl, 31);
        const bool bad_dist = is_match && dist > o;
        const bool over = have && (uint64_t)o + len > cap;
        const uint32_t fails = simt::ballot(bad_dist || over);
        if (fails) {
            const int f = simt::ffs(fails) - 1;
            pos_io = pos;
            return simt::shfl(over ? (uint32_t)ST_OUT_OVERFLOW : (uint32_t)ST_BAD_DISTANCE, f);
        }
        simt::syncwarp();  // the previous step's bytes are in place
        if (have && !is_match) out[o] = (uint8_t)t;
        const bool free_m = is_match && len <= 16 && o + len <= pos + dist;
        if (free_m) {
            const uint8_t *src = out + (o - dist);
            uint8_t *dst = out + o;
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = (uint32_t)j < len ? src[j] : 0u;
#pragma unroll
            for (int j = 0; j < 16; j++)
                if ((uint32_t)j < len) dst[j] = (uint8_t)v[j];
        }
        uint32_t m = simt::ballot(is_match && !free_m);
        while (m) {
            const int j = simt::ffs(m) - 1;
            m &= m - 1;
            copy_match_slow(out, simt::shfl(o, j), simt::shfl(len, j), simt::shfl(dist, j));  // (syncs the warp first)
        }
        pos += total;
    }
    simt::syncwarp();
    pos_io = pos;
    return ST_OK;
}

// The block loop of inflate() (inflate.c:896-1950): decodes blocks from the window position, which
// must be a block header, until the final block has ended (BLK_FINAL), rule Q2 ended the stream
// (BLK_Q2), or a block boundary at or past ring-coordinate bit `stop_bit` has been reached (BLK_STOP,
// used by the block-split path; ~0 = never).
enum { BLK_FINAL = 0, BLK_Q2 = 1, BLK_STOP = 2 };
template <int SINK>
DBG_DEV uint32_t inflate_blocks(Window &w, const StreamIn &g, InflateSmem *sm, Sink &k, uint64_t stop_bit, uint32_t &end)
{
    BlockTables bt;
    bt.lit_max = bt.dist_max = 0;
    end = BLK_FINAL;
    for (;;) {
        if (w.abs_bits() >= stop_bit) {
            end = BLK_STOP;
            return ST_OK;
        }
        if (w.abs_bits() >= 8 * g.end_byte) return ST_TRUNCATED;
        const uint32_t hdr = w.peek32() & 7;
        w.consume(3);
        const bool more = !(hdr & 1);
        const uint32_t btype = hdr >> 1;
        if (btype == 0) {
            w.consume((0u - w.s) & 7);  // to the next byte boundary (32*wb is byte aligned)
            const uint32_t v = w.peek32();
            const uint32_t len = v & 0xffff, nlen = v >> 16;
            w.consume(32);
            if (len != (~nlen & 0xffff)) return ST_STORED_LEN;
            if (len) {
                const uint64_t bytepos = w.abs_bits() >> 3;
                if (bytepos + len > g.end_byte) return ST_TRUNCATED;
                if (SINK == SINK_TOKENS) {
                    if (len > 256) {
                        // a sizeable stored block: one token per byte would make the expansion slower than decoding
                        // the chunk again (where stored bytes are a lane-parallel copy), so give up on tokens here
                        k.tok_cap = 0;
                        k.ntok = 0x40000000u;
                    } else {  // stored bytes become literal tokens
                        const uint8_t *src = w.base + bytepos;
                        for (uint32_t i = (uint32_t)simt::lane(); i < len; i += 32)
                            if (k.ntok + i < k.tok_cap) k.tok[k.ntok + i] = src[i];
                        k.ntok += len;
                    }
                } else if (SINK != SINK_COUNT) {
                    if ((uint64_t)k.pos + len > k.cap) return ST_OUT_OVERFLOW;
                    if (SINK == SINK_BYTES) {
                        copy_stored(k.out + k.pos, w.base + bytepos, len);
                    } else {
                        flush_pending16(k.pd);
                        copy_stored_u16(k.out16 + k.pos, w.base + bytepos, len);
                        simt::syncwarp();
                    }
                }
                k.pos += len;
                w.seek(bytepos + len);
            }
        } else if (btype != 3) {  // btype 3: inflate.c:990-998, ignored in the no-assert build
            uint32_t st = read_huffman_tables(w, sm, btype, bt);
            if (st) return st;
            uint32_t why = END_EOB;
            // Lane-parallel rounds (lane_round above) alternate with the lane-cooperative symbol walk (decode_symbols) until
            // the block has ended. SINK_TOKENS: the round's extent is the caller's next boundary hint and the tokens go
            // straight to the chunk's token area; where a round ends short of the block's end the walk decodes a stretch and another
            // round follows. SINK_BYTES: no hint, so the
            // block is taken in rounds of k.round_bits, each round's tokens go through the warp's scratch and are expanded
            // into the output right away; where a round's chain breaks the walk decodes a stretch and the rounds go on.
            bool block_done = false;
            bool use_lanes = (SINK == SINK_TOKENS || SINK == SINK_BYTES) && k.lanes;
            if (use_lanes && k.lane_skip) {
                k.lane_skip--;  // this block is left to the walk (rounds achieved nothing on the blocks before it)
                use_lanes = false;
            }
            for (;;) {
                bool stretch = false;
                if (use_lanes) {
                    for (;;) {
                        const uint64_t p0 = w.abs_bits();
                        const uint64_t in_end = 8 * g.end_byte;
                        LaneRound lr;
                        uint32_t got;
                        if (SINK == SINK_TOKENS) {
                            const bool hinted = stop_bit < in_end;
                            uint64_t ext = (hinted ? stop_bit : in_end) > p0 ? (hinted ? stop_bit : in_end) - p0 : 0;
                            if (!hinted && ext > LB_NOHINT_EXTENT) ext = LB_NOHINT_EXTENT;
                            if (k.ntok >= k.tok_cap) {
                                use_lanes = false;
                                break;
                            }
                            got = lane_round<true>(sm, bt, w.base, g, p0, ext, k.tok + k.ntok, k.tok_cap - k.ntok, lr, k.lb_stats);
                            if (got == LB_UNUSED) {
                                use_lanes = false;  // (too little left of the extent, or no room for tokens)
                                break;
                            }
                            k.ntok += lr.ntok;
                            k.pos += lr.out;
                        } else {
                            uint64_t ext = in_end > p0 ? in_end - p0 : 0;
                            if (ext > k.round_bits) ext = k.round_bits;
                            got = lane_round<false>(sm, bt, w.base, g, p0, ext, k.tok, k.tok_cap, lr, k.lb_stats);
                            if (got == LB_UNUSED) {
                                if (ext >= 4 * LB_MIN_SUB) {  // (not the short tail of a stream)
                                    k.lane_skip = k.lane_bad < 8 ? k.lane_bad : 8;
                                    k.lane_bad++;
                                    stretch = true;
                                } else {
                                    use_lanes = false;
                                }
                                break;
                            }
                            simt::syncwarp();
                            flush_pending(k.pd);
                            // the runs that count, one after the other, each from its own stretch of the scratch
                            for (uint32_t j = 0; j < lr.used; j++) {
                                st = expand_tokens_bytes(k.tok + (uint64_t)j * lr.stride, simt::shfl(lr.mine, (int)j), k.out, k.pos, k.cap);
                                if (st) return st;
                            }
                        }
                        w.seek_bits(lr.resume);
                        if (got == LB_DONE) {
                            block_done = true;
                            k.lane_bad = 0;
                            break;
                        }
                        if (SINK == SINK_TOKENS) {
                            // The chain broke before the end of the block (a lane without a merge point, a dense stretch), or the
                            // block is longer than the extent. The walk decodes a stretch and another round takes what is left
                            // -- leaving the rest of the block to the walk costs a lone stream milliseconds (gzipsample.gz: one
                            // of its five blocks). A chain that breaks in the first lanes is not tried again.
                            if (lr.used > 2) stretch = true;
                            else use_lanes = false;
                            break;
                        }
                        if (lr.resume - p0 < k.round_bits / 2) {  // little progress: the chain broke early
                            // Broken in the very first lanes: data whose chains do not merge (run-length trains, window-limit
                            // periods). Back off: after the n-th such round in a row the next min(n - 1, 8) blocks are left to
                            // the walk. A chain that breaks further out (a dense stretch, a lane that found no merge point)
                            // costs nothing to try again behind the spot.
                            if (lr.used <= 2) {
                                k.lane_skip = k.lane_bad < 8 ? k.lane_bad : 8;
                                k.lane_bad++;
                            }
                            stretch = true;
                            break;
                        }
                        k.lane_bad = 0;
                    }
                    if (block_done) break;
                    if (k.lane_skip) use_lanes = false;  // backing off: the walk takes the rest of this block
                }
                // the walk: to the end of the block, or -- between rounds -- over a stretch of a quarter round
                bool temp_limit = false;
                if ((SINK == SINK_BYTES || SINK == SINK_TOKENS) && use_lanes && stretch) {
                    const uint64_t lim = w.abs_bits() + (SINK == SINK_BYTES ? k.round_bits / 4 : LB_TOKENS_STRETCH);
                    if (lim < g.q2_limit) {
                        w.set_limit(lim);
                        temp_limit = true;
                    }
                }
                why = END_EOB;
                st = decode_symbols<SINK>(w, sm, bt, k, why);

                if (temp_limit) w.set_limit(g.q2_limit);
                if (st) return st;
                if (temp_limit && why == END_LIMIT) continue;  // back to the rounds
                break;
            }
            if (block_done) {
                if (!more) return ST_OK;
                continue;
            }
            if (why == END_LIMIT) {  // rule Q2
                end = BLK_Q2;
                return ST_OK;
            }
        }
        if (!more) return ST_OK;
    }
}

// Decodes one raw DEFLATE stream of `in_size` bytes at `in` into out[0..cap).
// Every lane of the warp must call it with identical arguments; the return
// value and *final_size are uniform. `in` may have any alignment; bytes from
// (in & ~15) up to the 16-byte boundary at or after in + in_size must be readable.
DBG_DEV uint32_t inflate_warp(InflateSmem *sm, const uint8_t *in, uint64_t in_size, uint8_t *out, uint64_t cap,
                              uint64_t *final_size, uint32_t *tok_scratch = nullptr, uint32_t *lb_stats = nullptr,
                              uint32_t round_bits = LB_ROUND_BITS)
{
    *final_size = 0;
    if (cap < in_size) return ST_CAP_LT_INPUT;
    if (in_size < 5) return ST_INPUT_TOO_SMALL;
    if (in_size >= (1ull << 31) || cap >= (1ull << 32) - 1024) return ST_TOO_LARGE;  // keeps pos + len in 32 bits

    Window w;
    const StreamIn g = open_stream(w, sm, in, in_size);
    w.seek(g.mis);

    Sink k;
    k.out = out;
    k.out16 = nullptr;
    k.abs_base = 0;
    k.pos = 0;
    k.cap = (uint32_t)cap;
    k.pd.ptr = out;
    k.pd.val = 0;
    k.pd.on = false;
    k.tok = tok_scratch;  // LB_ROUND_TOKENS slots of this warp's own: Huffman blocks are decoded in lane-parallel rounds
    k.ntok = 0;
    k.tok_cap = tok_scratch ? LB_ROUND_TOKENS : 0;
    k.lanes = tok_scratch != nullptr;
    k.lb_stats = lb_stats;
    k.round_bits = round_bits;
    k.lane_bad = k.lane_skip = 0;
    uint32_t end;
    const uint32_t st = inflate_blocks<SINK_BYTES>(w, g, sm, k, ~0ull, end);
    if (st) return st;
    flush_pending(k.pd);
    *final_size = k.pos;
    return ST_OK;
}

// ------------------------------------------------------- chunk-parallel paths --
// Streams that one warp would take too long for are cut into pieces that are decoded side by side into 16-bit cells
// (a value >= 256 is a marker "the byte at distance 32768 - (v - 256) before this piece's output"), which the resolve
// kernels of split_kernels.cuh turn into bytes: single fixed-Huffman blocks by fx_core.h (one LANE per chunk), long
// multi-block streams by bsplit_core.h (one warp per stretch between two block boundaries).
struct ChunkResult {
    uint64_t exit_bits;  // stream-relative bit where the next chunk's first symbol starts
    uint32_t out_bytes;
    uint32_t flag;
    uint32_t ntok;       // SINK_TOKENS: symbols seen (more than the token area holds = overflow)
};

// What stb_image_write emits: ONE final fixed-Huffman block (fx_core.h takes such streams).
DBG_DEV bool is_single_fixed_block(const uint8_t *in) { return (in[0] & 7) == 3; }  // BFINAL=1, BTYPE=01


}  // namespace dbg
