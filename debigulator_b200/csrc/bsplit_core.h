// bsplit_core.h -- block-split path: intra-stream parallelism for ordinary multi-block DEFLATE
// streams (what zlib / gzip write), for batches whose longest streams would otherwise keep one warp
// busy long after the rest of the GPU has drained (BASELINE config 5: members of up to 16 MiB).
//
// A stream is cut into regions of `region_bits` compressed bits. For every region but the first a
// warp SEARCHES the first bit position that looks like the header of a dynamic-Huffman block
// (BTYPE = 2, HLIT / HDIST in range, a complete code-length code, code lengths that describe complete
// literal/length and distance codes with an end-of-block symbol). Such a position is only a HINT:
//   count    one warp per hinted position decodes (sizes only) from it to the first block boundary at
//            or past the next hint;
//   chain    one thread per stream checks, from the real stream start, that every chunk ends exactly
//            on the next hint. By induction every chunk then started on a true block boundary and the
//            concatenation of the chunk decodes IS the sequential decode, bit for bit (same block
//            parser, same table builder, same rule-Q2 limit). A stream whose chain does not close is
//            handed back to the warp-per-stream kernel untouched;
//   decode   one warp per chunk decodes into 16-bit cells (markers for matches that reach before the
//            chunk), resolved by the same tail / body kernels as the split-stream path.
// Reference semantics are therefore those of inflate_blocks() in inflate_core.h; nothing here
// interprets the stream on its own authority.
#pragma once
#include "inflate_core.h"

namespace dbg {

constexpr uint32_t REGION_BYTES = 65536;
constexpr uint32_t BS_QCAP = 128;
constexpr uint64_t BS_NONE = ~0ull;

struct SearchSmem {        // per warp
    uint32_t q[BS_QCAP];   // bit offsets (relative to the search start) that passed the code-length-code test
    uint32_t qn;
    uint32_t pad[3];
};

// Word-granular view of a byte-addressed stream (any alignment); reads past the last word give zero.
struct BitSrc {
    const uint32_t *a;
    uint32_t boff;   // bit offset of the stream start inside a[0]
    uint64_t last;   // index of the last word that holds stream bytes
};
DBG_DEV BitSrc bit_src(const uint8_t *in, uint64_t in_size)
{
    BitSrc b;
    const uintptr_t p = (uintptr_t)in;
    b.a = (const uint32_t *)(p & ~(uintptr_t)3);
    b.boff = 8 * (uint32_t)(p & 3);
    b.last = (b.boff + 8 * in_size - 1) >> 5;
    return b;
}
DBG_DEV uint32_t bs_word(const BitSrc &b, uint64_t i) { return i <= b.last ? simt::ldg_u32(b.a + i) : 0u; }
DBG_DEV uint32_t bs_peek(const BitSrc &b, uint64_t abit)
{
    const uint64_t i = abit >> 5;
    return simt::funnel_r(bs_word(b, i), bs_word(b, i + 1), (uint32_t)abit & 31);
}

// kraft12[v] = sum over the four 3-bit code lengths packed in v of 2^(7-len) (0 for len 0).
DBG_DEV void build_kraft12(uint16_t *lut, uint32_t tid, uint32_t nthreads)
{
    for (uint32_t i = tid; i < 4096; i += nthreads) {
        uint32_t s = 0;
        for (int j = 0; j < 4; j++) {
            const uint32_t l = (i >> (3 * j)) & 7;
            s += l ? 128u >> l : 0u;
        }
        lut[i] = (uint16_t)s;
    }
}

// Lane-local check of the dynamic block header whose BFINAL bit sits at absolute bit `abit`:
// decodes the HLIT + HDIST code lengths with the code-length code and accepts only what a
// compressor writes -- the lengths fill exactly HLIT + HDIST entries, symbol 256 has a code, the
// literal/length code is complete, the distance code is complete or has at most one code.
DBG_DEV bool validate_dynamic_header(const BitSrc &b, uint64_t abit)
{
    const uint32_t h = bs_peek(b, abit);
    const uint32_t hlit = ((h >> 3) & 31) + 257, hdist = ((h >> 8) & 31) + 1, hclen = ((h >> 13) & 15) + 4;
    uint64_t at = abit + 17;
    // code-length-code lengths by symbol (3 bits each), counts per length (8 bits each)
    uint64_t pl = 0, cntp = 0;
    for (uint32_t i = 0; i < hclen; i += 8) {
        const uint32_t v = bs_peek(b, at + 3 * i);
        for (uint32_t j = 0; j < 8 && i + j < hclen; j++) {
            const uint64_t l = (v >> (3 * j)) & 7;
            pl |= l << (3 * swizzle_at(i + j));
            if (l) cntp += 1ull << (8 * l);
        }
    }
    at += 3 * hclen;
    // symbols sorted by (length, symbol): 19 x 5 bits in s0 (first 12) and s1
    uint64_t offp = 0, s0 = 0, s1 = 0;
    {
        uint64_t o = 0;
        for (uint32_t l = 1; l < 8; l++) {
            offp |= o << (8 * l);
            o += (cntp >> (8 * l)) & 255;
        }
        for (uint64_t sym = 0; sym < 19; sym++) {
            const uint32_t l = (uint32_t)(pl >> (3 * sym)) & 7;
            if (!l) continue;
            const uint32_t idx = (uint32_t)(offp >> (8 * l)) & 255;
            offp += 1ull << (8 * l);
            if (idx < 12) s0 |= sym << (5 * idx);
            else s1 |= sym << (5 * (idx - 12));
        }
    }
    const uint32_t n = hlit + hdist;
    uint32_t i = 0, prev = 0, kraft_lit = 0, kraft_dist = 0, ndist = 0, len256 = 0;
    while (i < n) {
        const uint32_t bits = bs_peek(b, at);
        // canonical decode, one bit at a time (at most 7)
        uint32_t code = 0, first = 0, index = 0, sym = 99, used = 0;
        for (uint32_t l = 1; l < 8; l++) {
            code |= (bits >> (l - 1)) & 1;
            const uint32_t count = (uint32_t)(cntp >> (8 * l)) & 255;
            if (code < first + count) {
                const uint32_t k = index + code - first;
                sym = (uint32_t)((k < 12 ? s0 >> (5 * k) : s1 >> (5 * (k - 12))) & 31);
                used = l;
                break;
            }
            index += count;
            first = (first + count) << 1;
            code <<= 1;
        }
        if (sym == 99) return false;
        uint32_t rep = 1, val = sym;
        if (sym == 16) {
            if (i == 0) return false;
            rep = 3 + ((bits >> used) & 3);
            used += 2;
            val = prev;
        } else if (sym == 17) {
            rep = 3 + ((bits >> used) & 7);
            used += 3;
            val = 0;
        } else if (sym == 18) {
            rep = 11 + ((bits >> used) & 127);
            used += 7;
            val = 0;
        }
        at += used;
        if (i + rep > n) return false;
        if (val) {
            const uint32_t nl = i < hlit ? (rep < hlit - i ? rep : hlit - i) : 0;
            kraft_lit += nl * (32768u >> val);
            kraft_dist += (rep - nl) * (32768u >> val);
            ndist += rep - nl;
        }
        if (i <= 256 && 256 < i + rep) len256 = val;
        prev = val;
        i += rep;
    }
    return len256 != 0 && kraft_lit == 32768u && (kraft_dist == 32768u || ndist <= 1);
}

// First plausible dynamic-block header at a stream bit in [lo_bit, hi_bit), or BS_NONE. Warp-wide,
// uniform result. Every lane tests the 32 bit positions of one input word per step.
DBG_DEV uint64_t find_block_start(SearchSmem *q, const uint16_t *kraft12, const uint8_t *in, uint64_t in_size, uint64_t lo_bit,
                                  uint64_t hi_bit)
{
    const uint32_t ln = (uint32_t)simt::lane();
    const BitSrc b = bit_src(in, in_size);
    if (hi_bit > 8 * in_size) hi_bit = 8 * in_size;
    if (lo_bit >= hi_bit) return BS_NONE;
    const uint64_t a_lo = lo_bit + b.boff, a_hi = hi_bit + b.boff;
    const uint64_t w_first = a_lo >> 5, w_end = (a_hi + 31) >> 5;
    uint32_t best = 0xffffffffu;
    if (ln == 0) q->qn = 0;
    simt::syncwarp();
    for (uint64_t wbase = w_first; wbase < w_end; wbase += 32) {
        const uint64_t wi = wbase + ln;
        if (wi < w_end) {
            const uint32_t w0 = bs_word(b, wi), w1 = bs_word(b, wi + 1), w2 = bs_word(b, wi + 2), w3 = bs_word(b, wi + 3);
            const uint64_t X = (uint64_t)w0 | ((uint64_t)w1 << 32);
            // BTYPE == 2, HLIT <= 29, HDIST <= 29 for all 32 positions at once
            uint32_t m = (uint32_t)(~(X >> 1) & (X >> 2) & ~((X >> 4) & (X >> 5) & (X >> 6) & (X >> 7)) &
                                    ~((X >> 9) & (X >> 10) & (X >> 11) & (X >> 12)));
            const uint64_t p0 = wi << 5;
            if (p0 < a_lo) m &= ~0u << (uint32_t)(a_lo - p0);
            if (p0 + 32 > a_hi) m &= (1u << (uint32_t)(a_hi - p0)) - 1;  // a_hi - p0 is 1..31 here
            while (m) {
                const uint32_t k = (uint32_t)simt::ffs(m) - 1;
                m &= m - 1;
                const uint32_t hclen = ((uint32_t)(X >> (k + 13)) & 15) + 4;
                const uint32_t t = k + 17;  // the code-length-code lengths start here
                const bool low = t < 32;
                const uint32_t c0 = low ? w0 : w1, c1 = low ? w1 : w2, c2 = low ? w2 : w3;
                uint64_t V = (uint64_t)simt::funnel_r(c0, c1, t & 31) | ((uint64_t)simt::funnel_r(c1, c2, t & 31) << 32);
                V &= (1ull << (3 * hclen)) - 1;
                const uint32_t sum = kraft12[V & 4095] + kraft12[(V >> 12) & 4095] + kraft12[(V >> 24) & 4095] +
                                     kraft12[(V >> 36) & 4095] + kraft12[(V >> 48) & 4095];
                if (sum == 128) {
                    const uint32_t idx = simt::atomic_inc_shared(&q->qn);
                    if (idx < BS_QCAP) q->q[idx] = (uint32_t)(p0 + k - a_lo);
                }
            }
        }
        simt::syncwarp();
        uint32_t qn = q->qn;
        if (qn > BS_QCAP) qn = BS_QCAP;
        const bool last = wbase + 32 >= w_end;
        if (qn >= 32 || (last && qn)) {
            for (uint32_t base = 0; base < qn; base += 32) {
                const uint32_t idx = base + ln;
                uint32_t rel = 0xffffffffu;
                if (idx < qn) {
                    const uint32_t r = q->q[idx];
                    if (validate_dynamic_header(b, a_lo + r)) rel = r;
                }
                for (int d = 16; d; d >>= 1) {
                    const uint32_t o = simt::shfl_xor(rel, d);
                    rel = o < rel ? o : rel;
                }
                best = rel < best ? rel : best;
            }
            simt::syncwarp();
            if (ln == 0) q->qn = 0;
            simt::syncwarp();
            if (best != 0xffffffffu) break;
        }
    }
    return best == 0xffffffffu ? BS_NONE : lo_bit + best;
}

// Decodes blocks from stream bit `start_bit` (a block header) to the first block boundary at or
// past `stop_bit` (BS_NONE: to the end of the stream). SINK_COUNT: sizes only; SINK_U16: into cells.
template <int SINK>
DBG_DEV ChunkResult decode_block_chunk(InflateSmem *sm, const uint8_t *in, uint64_t in_size, uint64_t start_bit, uint64_t stop_bit,
                                       uint16_t *cells, uint32_t cell_cap, uint64_t abs_base, uint32_t *tok = nullptr,
                                       uint32_t tok_cap = 0, bool lanes = false, uint32_t *lb_stats = nullptr)
{
    ChunkResult r;
    Window w;
    const StreamIn g = open_stream(w, sm, in, in_size);
    const uint64_t off = 8ull * g.mis;
    w.seek_bits(start_bit + off);
    Sink k;
    k.out = nullptr;
    k.out16 = cells;
    k.abs_base = abs_base;
    k.pos = 0;
    k.cap = cell_cap;
    k.pd.ptr = nullptr;
    k.pd.val = 0;
    k.pd.on = false;
    k.tok = tok;
    k.ntok = 0;
    k.tok_cap = tok_cap;
    k.lanes = lanes;
    k.lb_stats = lb_stats;
    k.round_bits = LB_ROUND_BITS;
    k.lane_bad = k.lane_skip = 0;
    uint32_t end = BLK_FINAL;
    const uint32_t st = inflate_blocks<SINK>(w, g, sm, k, stop_bit == BS_NONE ? BS_NONE : stop_bit + off, end);
    if (SINK == SINK_U16) flush_pending16(k.pd);
    r.exit_bits = w.abs_bits() - off;
    r.out_bytes = k.pos;
    r.ntok = k.ntok;
    if (st) r.flag = CH_ERR + st;
    else r.flag = end == BLK_STOP ? CH_RUN : end == BLK_FINAL ? CH_EOB : CH_Q2;
    return r;
}

// Second pass over a chunk whose symbols were recorded as tokens: expands them into the chunk's 16-bit
// cells. 32 tokens per step: an exclusive scan of their lengths gives every token its output offset, all
// literals are stored at once, then the matches -- those that only read what lies before this step's
// output and are short are copied by their own lane, all at the same time; the others (they may read
// what this step writes, or are long) one after the other with the whole warp. Returns ST_OK or the
// status of a distance that reaches before the start of the stream (inflate.c:1843).
DBG_DEV uint32_t expand_tokens_warp(const uint32_t *tok, uint32_t ntok, uint16_t *cells, uint64_t abs_base, uint32_t *out_bytes)
{
    const uint32_t ln = (uint32_t)simt::lane();
    uint32_t pos = 0, err = ST_OK;
    uint32_t t_next = ln < ntok ? simt::ldg_u32(tok + ln) : 0u;
    for (uint32_t base = 0; base < ntok; base += 32) {
        const bool have = base + ln < ntok;
        const uint32_t t = t_next;
        t_next = base + 32 + ln < ntok ? simt::ldg_u32(tok + base + 32 + ln) : 0u;  // the next step's tokens are on their way
        const bool is_match = have && (t & TOKEN_MATCH);
        const uint32_t len = !have ? 0u : is_match ? (t >> 16) & 0x1ff : 1u;
        const uint32_t dist = (t & 0x7fff) + 1;
        uint32_t incl = len;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = simt::shfl_up(incl, d);
            if (ln >= (uint32_t)d) incl += y;
        }
        const uint32_t o = pos + incl - len;        // this token's output offset
        const uint32_t total = simt::shfl(incl, 31);
        if (have && !is_match) cells[o] = (uint16_t)t;
        // a match is "free" when its whole source lies before this step's output and it is short
        const bool bad = is_match && dist > abs_base + o;
        const bool free_m = is_match && !bad && len <= 24 && o + len <= pos + dist;
        if (simt::any(bad)) {
            err = ST_BAD_DISTANCE;
            break;
        }
        if (free_m) {
            // nothing this step writes is read here, so the loads of a round all go out before its stores (a
            // load-store-load chain would pay one L2 round trip per cell)
            const int32_t s0 = (int32_t)o - (int32_t)dist;
            if (s0 >= 0) {
                const uint16_t *src = cells + s0;
                uint16_t *dst = cells + o;
                for (uint32_t rem = len; ; rem -= 8) {
                    uint32_t v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = (uint32_t)j < rem ? src[j] : 0u;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if ((uint32_t)j < rem) dst[j] = (uint16_t)v[j];
                    if (rem <= 8) break;
                    src += 8;
                    dst += 8;
                }
            } else {  // reaches before the group: markers for that part
                for (uint32_t i = 0; i < len; i++) {
                    const int32_t si = s0 + (int32_t)i;
                    const uint32_t v = si < 0 ? (uint32_t)(256 + 32768 + si) : cells[si];
                    cells[o + i] = (uint16_t)v;
                }
            }
        }
        simt::syncwarp();  // literals and free matches of this step are in place
        uint32_t m = simt::ballot(is_match && !free_m);
        while (m) {
            const int j = simt::ffs(m) - 1;
            m &= m - 1;
            copy_match_u16(cells, simt::shfl(o, j), simt::shfl(len, j), simt::shfl(dist, j));
        }
        pos += total;
    }
    *out_bytes = pos;
    return err;
}

// A lone long stream has few chunks (one per block boundary the search found), and the expansion of a chunk is one warp's
// latency chain (32 tokens per step, one L2 round trip each): gzipsample.gz spends 2.3 of its 3.9 ms there. A chunk's token
// run can be cut anywhere: piece j starts at token j * piece_tok and at the output offset of everything before it, which is
// a prefix sum of the token lengths. Every piece is then a marker domain of its own (a match that reaches before the piece
// becomes a marker, resolved against the finished bytes of the pieces before it), expanded by a warp of its own.
// Writes p_off[j] (output offset inside the stream) and p_len[j] for j < ceil(ntok / piece_tok); piece_tok must be a
// multiple of 128 (the walk takes four tokens per lane and step, all four loads in flight at once). Warp-wide, uniform
// result = the number of pieces.
DBG_DEV uint32_t cut_token_pieces(const uint32_t *tok, uint32_t ntok, uint32_t piece_tok, uint64_t out0, uint64_t *p_off, uint32_t *p_len)
{
    const uint32_t ln = (uint32_t)simt::lane();
    uint32_t run = 0, start = 0, j = 0;
    for (uint32_t base = 0; base < ntok; base += 128) {
        if (base % piece_tok == 0) {
            if (ln == 0) {
                p_off[j] = out0 + run;
                if (j) p_len[j - 1] = run - start;
            }
            start = run;
            j++;
        }
        uint32_t t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) t[k] = base + 32 * k + ln < ntok ? simt::ldg_u32(tok + base + 32 * k + ln) : 0u;
        uint32_t len = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) len += base + 32 * k + ln >= ntok ? 0u : (t[k] & TOKEN_MATCH) ? (t[k] >> 16) & 0x1ff : 1u;
        for (int d = 16; d; d >>= 1) len += simt::shfl_xor(len, d);
        run += len;
    }
    if (j && ln == 0) p_len[j - 1] = run - start;
    simt::syncwarp();
    return j;
}

}  // namespace dbg
