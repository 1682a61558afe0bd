// fx_core.h -- lane-serial decode of single fixed-Huffman-block DEFLATE streams (device code, sm_100a).
//
// What stb_image_write emits for every PNG is ONE final fixed-Huffman block (stb_write.h:913-916): BASELINE
// configs 3 and 4. The reference decodes such a stream with the generic loop of inflate() (inflate.c:1018-1181
// for the table, :1697-1909 for the symbols). The warp-per-stream decoder of inflate_core.h spends ~64 warp
// instructions per symbol on it; this file spends ~2 by giving every LANE its own piece of the stream:
//
//   head      one warp per chunk of `chunk_bytes` compressed bytes. A fixed-code symbol is at most 31 bits
//             long, so one of the 32 bit offsets at the start of the chunk is a true symbol start. Lane j
//             decodes (sizes only) from offset j for FX_HEAD_BITS bits. Fixed-Huffman chains started at
//             neighbouring offsets merge quickly (measured on stb streams: 1.3 distinct chains left after
//             256 bits, 1.00 after 1,024 bits in every one of 1,100 chunks), so what is left is a handful
//             of SURVIVOR positions, normally one. Every hypothesis j records where it stopped and which
//             survivor that is. Nothing is guessed: all 32 hypotheses are kept, merged or not.
//   sizes     one LANE per survivor decodes from its position across the end of its chunk into the head
//             region of the next one, until it stands exactly where the hypothesis it entered on stopped:
//             that is the next chunk's survivor it continues as. It records bytes, symbols, exit.
//   chain     per stream, from the one real start (bit 3 of chunk 0) along those links: which survivor of
//             every chunk is the real one, exact output offset and token offset of every chunk.
//   tokens    one lane per chunk decodes its real survivor again and writes one 32-bit token per symbol.
//   expand    one warp per GROUP of chunks turns tokens into 16-bit cells (expand_tokens_warp of
//             bsplit_core.h: 32 tokens per step); a match that reaches before the group becomes a marker.
//   resolve   the tail / body kernels of split_kernels.cuh: cells -> bytes.
//
// The concatenation of the real survivors' symbol runs is the sequential decode bit for bit: each run
// starts where the previous one ended, and all of them use the one step function fx_step below, which
// applies the reference's rules (end-of-stream rule Q2 inflate.c:1702-1717 before every symbol, litlen
// 286/287 and distance 30/31 fail, :1809).
#pragma once
#include "inflate_core.h"

namespace dbg {

constexpr uint32_t FX_HEAD_BITS = 1024;   // length of the merge run at the start of every chunk
constexpr uint32_t FX_CATCHUP = 48;       // extra single-symbol steps that let lagging chains reach the leading one
constexpr uint32_t FX_MIN_CHUNK = 2048;   // bytes; must exceed (FX_HEAD_BITS + 31 * (FX_CATCHUP + 2)) / 8

struct FxLuts {
    uint32_t lit[512];  // first 9 stream bits -> make_entry<K_LITLEN> of the fixed code (inflate.c:1035-1084)
    uint32_t dist[32];  // 5 stream bits -> make_entry<K_DIST> (5-bit codes read bit-reversed, :1783-1788)
};

DBG_DEV void fx_build_luts(FxLuts *L, uint32_t tid, uint32_t nthreads)
{
    for (uint32_t i = tid; i < 512; i += nthreads) {
        const uint32_t r9 = simt::brev(i) >> 23;  // the code as written MSB-first
        const uint32_t c7 = r9 >> 2, c8 = r9 >> 1;
        uint32_t sym, l;
        if (c7 < 24) { sym = 256 + c7; l = 7; }                // 0000000 .. 0010111
        else if (c8 < 192) { sym = c8 - 48; l = 8; }           // 00110000 .. 10111111
        else if (c8 < 200) { sym = 280 + (c8 - 192); l = 8; }  // 11000000 .. 11000111
        else { sym = 144 + (r9 - 400); l = 9; }                // 110010000 .. 111111111
        L->lit[i] = make_entry<K_LITLEN>(sym, l);
    }
    for (uint32_t i = tid; i < 32; i += nthreads) L->dist[i] = make_entry<K_DIST>(simt::brev(i) >> 27, 5);
}

// One symbol at the reader's position. Returns its kind; *nbits = code + extra bits, *len = output bytes,
// *tok = its token (TOKEN_MATCH | length << 16 | distance - 1, or the literal byte).
enum : uint32_t { FXK_LIT = 0, FXK_MATCH = 1, FXK_EOB = 2, FXK_BAD = 3 };
DBG_DEV uint32_t fx_symbol(const FxLuts *L, uint64_t buf, uint32_t *nbits, uint32_t *len, uint32_t *tok)
{
    const uint32_t x = (uint32_t)buf;
    const uint32_t e = L->lit[x & 511];
    const uint32_t l1 = e & 15;
    if (e & E_LIT) {
        *nbits = l1;
        *len = 1;
        *tok = e >> 16;
        return FXK_LIT;
    }
    if (e & E_BASE) {
        const uint32_t xb = (e >> 8) & 31;
        const uint32_t ln = (e >> 16) + ((x >> l1) & ((1u << xb) - 1));
        const uint32_t t1 = l1 + xb;                 // <= 8 + 5
        const uint32_t v = (uint32_t)(buf >> t1);
        const uint32_t e2 = L->dist[v & 31];
        if (!(e2 & E_BASE)) return FXK_BAD;          // distance symbols 30 / 31 (inflate.c:1809)
        const uint32_t xb2 = (e2 >> 8) & 31;
        const uint32_t dist = (e2 >> 16) + ((v >> 5) & ((1u << xb2) - 1));
        *nbits = t1 + 5 + xb2;                       // <= 31
        *len = ln;
        *tok = TOKEN_MATCH | (ln << 16) | (dist - 1);
        return FXK_MATCH;
    }
    *nbits = l1;
    *len = 0;
    *tok = 0;
    return (e & E_EOB) ? FXK_EOB : FXK_BAD;          // litlen 286 / 287
}

// Per-lane bit reader of the lane kernels below: like LaneBits (inflate_core.h), but it fetches 16 bytes at a time, two
// fetches ahead. Every lane reads a chunk of its own, so a 4-byte load moves a whole 32-byte sector per lane and relies on L1
// to keep it for the lane's next seven loads; with ~1,400 lanes per SM each on a line of its own it does not (ncu, round 2:
// L1 hit rate 26 %, 40 % of the lane kernels' stall samples on these loads). Same readable range as every other reader:
// from (address & ~15) to the 16-byte boundary at or after the stream end.
struct LaneBits16 {
    const uint32_t *a;  // 16-byte aligned address at or below the stream start
    uint32_t boff;      // bit offset of the stream start inside a[0..4)
    uint32_t last;      // index of the last readable 16-byte block; blocks past it read as zero
    uint32_t bidx;      // index of the block held in n0..n3
    uint32_t c0, c1, c2, c3, left;  // words of the current block that are not in `buf` yet (c0 first) and how many
    uint32_t n0, n1, n2, n3;        // the next block, on its way
    uint32_t nb;        // valid bits in buf (>= 32 whenever a symbol is decoded)
    uint64_t buf;       // next stream bits, LSB first

    DBG_DEVM void block(uint32_t i, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3) const
    {
        w0 = w1 = w2 = w3 = 0;
        if (i <= last) simt::ldg_u32x4(a + 4 * (uint64_t)i, w0, w1, w2, w3);
    }
    DBG_DEVM void open(const uint8_t *in, uint64_t in_size)
    {
        const uintptr_t p = (uintptr_t)in;
        a = (const uint32_t *)(p & ~(uintptr_t)15);
        boff = 8 * (uint32_t)(p & 15);
        last = (uint32_t)((((p + in_size + 15) & ~(uintptr_t)15) - (p & ~(uintptr_t)15)) >> 4) - 1;
    }
    DBG_DEVM uint32_t pop()
    {
        if (left == 0) {
            c0 = n0, c1 = n1, c2 = n2, c3 = n3;
            left = 4;
            bidx++;
            block(bidx, n0, n1, n2, n3);
        }
        const uint32_t r = c0;
        c0 = c1, c1 = c2, c2 = c3;
        left--;
        return r;
    }
    DBG_DEVM void seek(uint64_t stream_bit)
    {
        const uint64_t abit = stream_bit + boff;
        const uint32_t w = (uint32_t)(abit >> 5), sh = (uint32_t)abit & 31;
        bidx = (w >> 2) + 1;
        block(bidx - 1, c0, c1, c2, c3);
        block(bidx, n0, n1, n2, n3);
        left = 4;
        for (uint32_t k = 0; k < (w & 3); k++) pop();
        const uint64_t lo = pop();
        const uint64_t hi = pop();
        buf = (lo | (hi << 32)) >> sh;
        nb = 64 - sh;
    }
    DBG_DEVM void refill()
    {
        if (nb <= 32) {
            buf |= (uint64_t)pop() << nb;
            nb += 32;
        }
    }
    DBG_DEVM void drop(uint32_t n)
    {
        buf >>= n;
        nb -= n;
    }
};

// State of one decode run; positions are bits relative to the start of the run's chunk.
struct FxRun {
    uint32_t rel;       // next symbol starts here
    uint32_t out;       // bytes produced
    uint32_t ntok;      // symbols seen (end-of-block not counted)
    uint32_t flag;      // CH_RUN while running, else CH_EOB / CH_Q2 / CH_ERR + status
};

// Decodes symbols while rel < stop. `q2r` is the rule-Q2 limit relative to the chunk (no symbol may start at
// or past it). EMIT hands the tokens to the lane's writer.
template <bool EMIT, class Reader>
DBG_DEV void fx_run(const FxLuts *L, Reader &br, FxRun &r, uint32_t stop, uint32_t q2r, TokOut *tok)
{
    while (r.flag == CH_RUN && r.rel < stop) {
        if (r.rel >= q2r) {
            r.flag = CH_Q2;
            break;
        }
        br.refill();
        uint32_t nbits, len, t;
        const uint32_t kind = fx_symbol(L, br.buf, &nbits, &len, &t);
        if (kind == FXK_BAD) {
            r.flag = CH_ERR + ST_BAD_SYMBOL;
            break;
        }
        br.drop(nbits);
        r.rel += nbits;
        if (kind == FXK_EOB) {
            r.flag = CH_EOB;
            break;
        }
        if (EMIT) tok->put(t);
        r.ntok++;
        r.out += len;
    }
}

// Hypothesis record of (chunk, entry offset j): where the head run from offset j stopped and which survivor
// of the chunk that is. 0 = the hypothesis ended inside the head region (end-of-block, rule Q2, bad symbol):
// a run that enters the chunk on it ends the same way on its own.
constexpr uint32_t FX_HYP_LIVE = 0x80000000u;
DBG_DEV uint32_t fx_hyp_pack(uint32_t stop_rel, uint32_t surv) { return FX_HYP_LIVE | (surv << 24) | stop_rel; }
DBG_DEV uint32_t fx_hyp_stop(uint32_t h) { return h & 0xffffffu; }
DBG_DEV uint32_t fx_hyp_surv(uint32_t h) { return (h >> 24) & 31; }

DBG_DEV uint32_t fx_q2_rel(uint64_t in_size, uint64_t chunk_start_bit)
{
    const uint64_t q2 = 8 * in_size - 7;  // inflate.c:1702-1717: stop once ceil(P/8) >= size
    if (q2 <= chunk_start_bit) return 0;
    const uint64_t d = q2 - chunk_start_bit;
    return d > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)d;
}

// Head pass of chunk `c` (warp-wide). Writes hyp[0..32) and surv_start[0..nsurv); returns nsurv (uniform).
DBG_DEV uint32_t fx_head_warp(const FxLuts *L, const uint8_t *in, uint64_t in_size, uint32_t c, uint32_t chunk_bytes,
                              uint32_t *hyp, uint32_t *surv_start)
{
    const uint32_t ln = (uint32_t)simt::lane();
    const uint64_t cs = (uint64_t)c * chunk_bytes * 8;
    FxRun r;
    r.out = 0;
    r.ntok = 0;
    r.flag = CH_RUN;
    if (c == 0) {
        r.rel = 3;  // the one real entry: right behind BFINAL / BTYPE
    } else {
        const uint32_t q2r = fx_q2_rel(in_size, cs);
        LaneBits br;  // (measured: the 16-byte reader makes the head and sizes passes slower, 13.4 -> 15.9 ms on config 3)
        br.open(in, in_size);
        r.rel = ln;
        br.seek(cs + ln);
        fx_run<false>(L, br, r, FX_HEAD_BITS, q2r, nullptr);
        // let the chains that lag behind walk up to the leading one: most of the remaining merges happen here
        for (uint32_t it = 0; it < FX_CATCHUP; it++) {
            uint32_t lead = r.flag == CH_RUN ? r.rel : 0u;
            for (int d = 16; d; d >>= 1) {
                const uint32_t o = simt::shfl_xor(lead, d);
                lead = o > lead ? o : lead;
            }
            const bool behind = r.flag == CH_RUN && r.rel < lead;
            if (!simt::any(behind)) break;
            if (behind) fx_run<false>(L, br, r, r.rel + 1, q2r, nullptr);  // exactly one symbol
        }
    }
    const bool live = r.flag == CH_RUN;
    const uint32_t key = live ? r.rel : 0xffffffe0u + ln;  // ended hypotheses never group
    const uint32_t same = simt::match_any(key);
    const uint32_t leader = (uint32_t)simt::ffs(same) - 1;
    const uint32_t leaders = simt::ballot(live && leader == ln);
    const uint32_t surv = (uint32_t)simt::popc(leaders & ((1u << leader) - 1));
    hyp[ln] = live ? fx_hyp_pack(r.rel, surv) : 0u;
    if (live && leader == ln) surv_start[surv] = r.rel;
    return (uint32_t)simt::popc(leaders);
}

// Result of one survivor's sizes run.
struct alignas(16) FxRec {
    uint32_t out_bytes;
    uint32_t ntok;
    uint32_t exit_rel;  // where the run ended, relative to its chunk start
    uint32_t link;      // flag | next survivor index << 8 (CH_RUN only)
};

// Sizes run of the survivor that starts at `start_rel` of chunk c (one lane; no collectives).
// `next_hyp` = the 32 hypothesis records of chunk c + 1 (unused for the last chunk).
DBG_DEV FxRec fx_sizes_lane(const FxLuts *L, const uint8_t *in, uint64_t in_size, uint32_t c, uint32_t chunk_bytes,
                            uint32_t start_rel, const uint32_t *next_hyp)
{
    const uint64_t cs = (uint64_t)c * chunk_bytes * 8;
    const uint32_t chunk_bits = chunk_bytes * 8;
    const uint32_t q2r = fx_q2_rel(in_size, cs);
    LaneBits br;  // (measured: the 16-byte reader makes the head and sizes passes slower, 13.4 -> 15.9 ms on config 3)
    br.open(in, in_size);
    br.seek(cs + start_rel);
    FxRun r;
    r.rel = start_rel;
    r.out = 0;
    r.ntok = 0;
    r.flag = CH_RUN;
    fx_run<false>(L, br, r, chunk_bits, q2r, nullptr);
    if (r.flag == CH_RUN && r.rel >= q2r) r.flag = CH_Q2;  // the next symbol would start past the limit (last chunk: no chunk c + 1)
    uint32_t next = 0;
    if (r.flag == CH_RUN) {
        // now inside chunk c + 1, on entry offset rel - chunk_bits: walk to where that hypothesis stopped
        const uint32_t h = next_hyp[r.rel - chunk_bits];
        if (h & FX_HYP_LIVE) {
            const uint32_t target = chunk_bits + fx_hyp_stop(h);
            fx_run<false>(L, br, r, target, q2r, nullptr);
            if (r.flag == CH_RUN && r.rel != target) r.flag = CH_ERR + ST_BAD_CODE;  // cannot happen: same bits, same steps
            next = fx_hyp_surv(h);
        } else {
            // the hypothesis ended inside the head region, so this run ends there as well
            fx_run<false>(L, br, r, chunk_bits + FX_HEAD_BITS + 31 * (FX_CATCHUP + 2), q2r, nullptr);
            if (r.flag == CH_RUN) r.flag = CH_ERR + ST_BAD_CODE;  // cannot happen
        }
    }
    FxRec o;
    o.out_bytes = r.out;
    o.ntok = r.ntok;
    o.exit_rel = r.rel;
    o.link = r.flag | (next << 8);
    return o;
}

// Token run of a chunk's real survivor: the same symbols again, now written out. Returns the flag it ended
// with; *out_bytes / *ntok must equal what the sizes run recorded.
DBG_DEV uint32_t fx_tokens_lane(const FxLuts *L, const uint8_t *in, uint64_t in_size, uint32_t c, uint32_t chunk_bytes,
                                uint32_t start_rel, uint32_t exit_rel, uint32_t *tok, uint32_t *out_bytes, uint32_t *ntok)
{
    const uint64_t cs = (uint64_t)c * chunk_bytes * 8;
    const uint32_t q2r = fx_q2_rel(in_size, cs);
    LaneBits16 br;
    br.open(in, in_size);
    br.seek(cs + start_rel);
    FxRun r;
    r.rel = start_rel;
    r.out = 0;
    r.ntok = 0;
    r.flag = CH_RUN;
    TokOut w;
    w.open(tok);
    fx_run<true>(L, br, r, exit_rel, q2r, &w);
    w.close();
    *out_bytes = r.out;
    *ntok = r.ntok;
    return r.flag;
}

}  // namespace dbg
