// bsplit_kernels.cuh -- __global__ wrappers of the block-split path (bsplit_core.h): chunk-parallel
// decode of long multi-block DEFLATE streams. Shares the cell -> byte resolve kernels with the
// split-stream path (split_kernels.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bsplit_core.h"
#include "split_kernels.cuh"

namespace dbg {

constexpr uint64_t BS_MIN_BYTES = 65536;  // shorter streams never take this path (4 regions of the smallest size)
constexpr int BS_WARPS_PER_CTA = 4;

struct BsSummary {         // device -> host after classify, and again after chain
    uint32_t n_split;
    uint32_t total_regions;
    uint64_t total_in;     // compressed bytes of the whole batch
    uint64_t cells_used;   // exact number of 16-bit cells (bs_chain_kernel)
    uint32_t n_fallback;   // streams whose chain did not close
    uint32_t pad;
    uint64_t split_in;     // compressed bytes of the streams taking this path
    uint64_t tok_bytes;    // allocation cursor of the token areas, in compressed bytes (bs_assign_kernel)
    uint32_t n_pruned;     // streams taken off this path again because their hints were too sparse (bs_prune_kernel)
    uint32_t pad2;
};

struct BsBatch {
    const uint8_t *in_base;
    const uint64_t *in_off;
    const uint64_t *in_size;
    uint8_t *out_base;
    const uint64_t *out_off;
    const uint64_t *out_cap;
    uint64_t *out_size;
    uint32_t *status;
    const uint32_t *pre_status;  // optional
    const uint32_t *taken;       // optional: streams already handled by the split-stream path
    uint32_t n;
    uint32_t resident_warps;     // streams the warp-per-stream kernel keeps in flight
    uint64_t min_bytes;          // lower bound of the split threshold
    uint32_t factor_q;           // threshold = factor_q / 4 x (batch bytes / resident warps)
    uint32_t region_bytes;       // compressed bytes per region
    BsSummary *summary;
    uint32_t *flag;         // per stream: 1 = block-split path (stays set for a handed-back stream)
    uint32_t *redo;         // per stream: 1 = handed back to the warp-per-stream kernel (second pass)
    uint32_t *chunk_base;   // per stream: first region index
    uint32_t *nchunks;      // per stream: regions
    uint64_t *cell_base;    // per stream
    uint32_t *chunk_stream; // per region
    uint64_t *cand;         // per region: hinted block start (stream bit) or BS_NONE
    uint64_t *exit_bits;    // per region: where its decode ended (count pass)
    uint64_t *c_out_off;    // per region: output offset inside the stream
    uint32_t *c_out_len;    // per region
    uint32_t *c_flag;       // per region
    uint16_t *cells;
    // tokens recorded by the count pass (optional: tok == nullptr = decode twice)
    uint32_t *tok;
    uint64_t *tok_stream_base;  // per stream: first token slot
    uint32_t *c_ntok;           // per region: symbols seen by the count pass
    uint32_t tok_per_byte;      // token slots per compressed byte; a chunk that needs more is Huffman-decoded again
    uint32_t lanes;             // the count pass may decode blocks lane-parallel (lane_round, inflate_core.h)
    uint32_t dyn_all;           // every stream of >= min_bytes that opens with a dynamic block takes this path
    uint32_t *lb_stats;         // per context, never reset: rounds / rounds with end-of-block / rounds without of lane_round,
                                // [3] chunks whose tokens were expanded, [4] chunks decoded a second time instead
};

// Token area of the chunk that starts at stream bit `start` and ends at the next hint (or the stream end).
__device__ __forceinline__ uint64_t bs_tok_base(const BsBatch &b, uint32_t s, uint64_t start)
{
    return b.tok_stream_base[s] + (uint64_t)b.tok_per_byte * (start >> 3);
}
__device__ __forceinline__ uint32_t bs_tok_cap(const BsBatch &b, uint32_t s, uint64_t start, uint64_t next_hint)
{
    const uint64_t end_byte = next_hint == BS_NONE ? b.in_size[s] : next_hint >> 3;
    return (uint32_t)((uint64_t)b.tok_per_byte * (end_byte - (start >> 3)));
}

__global__ void bs_sum_kernel(BsBatch b)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t v = s < b.n ? b.in_size[s] : 0;
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd((unsigned long long *)&b.summary->total_in, (unsigned long long)v);
}

// A stream is split when one warp would need clearly longer for it than the whole batch needs when
// it is spread evenly over the resident warps: compressed size > 2 x (batch bytes / resident warps).
__global__ void bs_classify_kernel(BsBatch b)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n) return;
    const uint64_t size = b.in_size[s], cap = b.out_cap[s];
    uint64_t thr = (b.summary->total_in / b.resident_warps) * b.factor_q / 4;
    if (thr < b.min_bytes) thr = b.min_bytes;
    // lane-parallel count pass: a stream that opens with a DYNAMIC block is worth this path whatever its size (its block
    // headers are found by the search, so every block gets its own warp and its own 32 lanes)
    if (b.dyn_all && size >= b.min_bytes && ((b.pre_status && b.pre_status[s]) ? 0u : ((b.in_base[b.in_off[s]] >> 1) & 3)) == 2) thr = b.min_bytes;
    uint32_t flag = 0;
    const bool ok = (!b.pre_status || b.pre_status[s] == 0) && (!b.taken || b.taken[s] == 0) && size >= thr && cap >= size &&
                    size < (1ull << 31) && cap < (1ull << 32) - 1024;
    // not worth it / not possible: a stream that opens with a stored block is (mostly) a plain copy, which
    // one warp does at ~0.8 GB/s and 16-bit cells would only slow down; a stream whose first block is also
    // its last has no block boundary to split at
    b.redo[s] = 0;
    const uint32_t first = ok ? b.in_base[b.in_off[s]] : 0u;
    if (ok && ((first >> 1) & 3) != 0 && (first & 1) == 0) {
        flag = 1;
        atomicAdd(&b.summary->n_split, 1u);
        atomicAdd((unsigned long long *)&b.summary->split_in, (unsigned long long)size);
    }
    b.flag[s] = flag;
}

// Regions per stream, once the host has chosen the region size (smaller regions when the split
// streams alone would not fill the GPU with 64 KiB ones).
__global__ void bs_assign_kernel(BsBatch b)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n || !b.flag[s]) return;
    const uint32_t nreg = (uint32_t)((b.in_size[s] + b.region_bytes - 1) / b.region_bytes);
    b.chunk_base[s] = atomicAdd(&b.summary->total_regions, nreg);
    b.nchunks[s] = nreg;
    if (b.tok)
        b.tok_stream_base[s] = (uint64_t)b.tok_per_byte *
                               atomicAdd((unsigned long long *)&b.summary->tok_bytes, (unsigned long long)b.in_size[s]);
}

__global__ void bs_fill_kernel(BsBatch b)
{
    const uint32_t s = blockIdx.x;
    if (!b.flag[s]) return;
    const uint32_t nch = b.nchunks[s], base = b.chunk_base[s];
    for (uint32_t c = threadIdx.x; c < nch; c += blockDim.x) b.chunk_stream[base + c] = s;
}

// Hints: one warp per region.
__global__ void __launch_bounds__(BS_WARPS_PER_CTA * 32) bs_search_kernel(BsBatch b)
{
    const uint32_t total_regions = b.summary->total_regions;
    __shared__ uint16_t kraft12[4096];
    __shared__ SearchSmem qs[BS_WARPS_PER_CTA];
    build_kraft12(kraft12, threadIdx.x, blockDim.x);
    __syncthreads();
    SearchSmem *q = &qs[threadIdx.x >> 5];
    const uint32_t warps = gridDim.x * BS_WARPS_PER_CTA;
    for (uint32_t t = blockIdx.x * BS_WARPS_PER_CTA + (threadIdx.x >> 5); t < total_regions; t += warps) {
        const uint32_t s = b.chunk_stream[t], c = t - b.chunk_base[s];
        uint64_t cand = 0;
        if (c) cand = find_block_start(q, kraft12, b.in_base + b.in_off[s], b.in_size[s], (uint64_t)c * b.region_bytes * 8,
                                       (uint64_t)(c + 1) * b.region_bytes * 8);
        if (simt::lane() == 0) b.cand[t] = cand;
        simt::syncwarp();
    }
}

// Hints too sparse? When the largest stretch between two consecutive chunk starts exceeds a third of the
// stream, the two passes of this path take longer than the single pass of one warp (a stream of long stored or
// fixed-Huffman stretches has no dynamic-block headers to find there). Such a stream is handed back before the
// count pass: it joins the streams of the second warp-per-stream pass. One thread per stream.
__global__ void bs_prune_kernel(BsBatch b)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n || !b.flag[s]) return;
    const uint32_t nch = b.nchunks[s], base = b.chunk_base[s];
    const uint64_t bits = 8 * b.in_size[s];
    uint64_t prev = 0, worst = 0;
    for (uint32_t c = 1; c < nch; c++) {
        const uint64_t cand = b.cand[base + c];
        if (cand == BS_NONE) continue;
        if (cand - prev > worst) worst = cand - prev;
        prev = cand;
    }
    if (bits - prev > worst) worst = bits - prev;
    if (worst > bits / 3) {
        for (uint32_t c = 0; c < nch; c++) b.cand[base + c] = BS_NONE;
        b.redo[s] = 1;  // flag[s] stays set: the first warp-per-stream pass is already running beside us
        atomicAdd(&b.summary->n_pruned, 1u);
    }
}

__device__ __forceinline__ uint64_t bs_next_hint(const BsBatch &b, uint32_t s, uint32_t t)
{
    const uint32_t end = b.chunk_base[s] + b.nchunks[s];
    for (uint32_t u = t + 1; u < end; u++)
        if (b.cand[u] != BS_NONE) return b.cand[u];
    return BS_NONE;
}

// Sizes: one warp per hinted region decodes to the first block boundary at or past the next hint.
__global__ void __launch_bounds__(BS_WARPS_PER_CTA * 32) bs_count_kernel(BsBatch b)
{
    const uint32_t total_regions = b.summary->total_regions;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateSmem *sm = reinterpret_cast<InflateSmem *>(smem_raw) + (threadIdx.x >> 5);
    const uint32_t warps = gridDim.x * BS_WARPS_PER_CTA;
    for (uint32_t t = blockIdx.x * BS_WARPS_PER_CTA + (threadIdx.x >> 5); t < total_regions; t += warps) {
        const uint64_t start = b.cand[t];
        if (start == BS_NONE) continue;
        const uint32_t s = b.chunk_stream[t];
        const uint64_t next = bs_next_hint(b, s, t);
        ChunkResult r;
        if (b.tok)
            r = decode_block_chunk<SINK_TOKENS>(sm, b.in_base + b.in_off[s], b.in_size[s], start, next, nullptr, 0, 0,
                                                b.tok + bs_tok_base(b, s, start), bs_tok_cap(b, s, start, next), b.lanes != 0,
                                                b.lb_stats);
        else
            r = decode_block_chunk<SINK_COUNT>(sm, b.in_base + b.in_off[s], b.in_size[s], start, next, nullptr, 0, 0);
        if (simt::lane() == 0) {
            b.exit_bits[t] = r.exit_bits;
            b.c_out_len[t] = r.out_bytes;
            b.c_flag[t] = r.flag;
            if (b.tok) b.c_ntok[t] = r.ntok;
        }
        simt::syncwarp();
    }
}

// One thread per stream: does every chunk end exactly on the next hint?
__global__ void bs_chain_kernel(BsBatch b)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n || !b.flag[s] || b.redo[s]) return;
    const uint32_t nch = b.nchunks[s], base = b.chunk_base[s];
    const uint64_t cap = b.out_cap[s];
    uint64_t pos = 0, expected = 0;
    uint32_t st = ST_OK;
    bool ended = false, fail = false;
    for (uint32_t c = 0; c < nch; c++) {
        const uint32_t t = base + c;
        const uint64_t cand = b.cand[t];
        const uint32_t len = b.c_out_len[t], flag = b.c_flag[t];
        b.c_out_off[t] = pos;
        if (cand == BS_NONE || ended || fail) {
            b.c_flag[t] = CH_IDLE;
            b.c_out_len[t] = 0;
            continue;
        }
        if (cand < expected) {
            // the previous chunk stepped over this hint: it was no block boundary. That chunk went on to the first real
            // boundary behind it (bs_count_kernel decodes "to the first block boundary at or past the next hint"), so
            // the hint is simply dropped; the chain goes on with the chunk that starts where the previous one ended.
            b.c_flag[t] = CH_IDLE;
            b.c_out_len[t] = 0;
            continue;
        }
        if (cand != expected) {  // a gap: no chunk starts where the previous one ended
            fail = true;
            b.c_flag[t] = CH_IDLE;
            b.c_out_len[t] = 0;
            continue;
        }
        if (flag >= CH_ERR) {
            // the sequential decoder reports whichever comes first: the overflow or the error
            st = pos + len > cap ? (uint32_t)ST_OUT_OVERFLOW : flag - CH_ERR;
            ended = true;
            b.c_flag[t] = CH_IDLE;
            b.c_out_len[t] = 0;
            continue;
        }
        pos += len;
        if (pos > cap) {
            st = ST_OUT_OVERFLOW;
            ended = true;
        } else if (flag != CH_RUN) {
            ended = true;
        } else {
            expected = b.exit_bits[t];
        }
    }
    if (!ended) fail = true;  // cannot happen: the last hinted chunk runs to the end of the stream
    if (fail) {
        // hand the stream back to the warp-per-stream kernel
        for (uint32_t c = 0; c < nch; c++) {
            b.c_flag[base + c] = CH_IDLE;
            b.c_out_len[base + c] = 0;
        }
        b.redo[s] = 1;  // flag[s] stays set: the first warp-per-stream pass may already be running beside us
        b.cell_base[s] = 0;
        atomicAdd(&b.summary->n_fallback, 1u);
        return;
    }
    b.status[s] = st;
    b.out_size[s] = st == ST_OK ? pos : 0;
    b.cell_base[s] = st == ST_OK ? atomicAdd((unsigned long long *)&b.summary->cells_used, (unsigned long long)pos) : 0;
}

// Chunk decode into 16-bit cells: one warp per hinted region.
__global__ void __launch_bounds__(BS_WARPS_PER_CTA * 32) bs_decode_kernel(BsBatch b)
{
    const uint32_t total_regions = b.summary->total_regions;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateSmem *sm = reinterpret_cast<InflateSmem *>(smem_raw) + (threadIdx.x >> 5);
    const uint32_t warps = gridDim.x * BS_WARPS_PER_CTA;
    for (uint32_t t = blockIdx.x * BS_WARPS_PER_CTA + (threadIdx.x >> 5); t < total_regions; t += warps) {
        const uint32_t s = b.chunk_stream[t];
        if (!b.flag[s] || b.redo[s] || b.status[s] != ST_OK || b.c_flag[t] == CH_IDLE) continue;
        const uint64_t stop = b.c_flag[t] == CH_RUN ? b.exit_bits[t] : BS_NONE;
        uint16_t *cells = b.cells + b.cell_base[s] + b.c_out_off[t];
        ChunkResult r;
        const uint64_t next = b.tok ? bs_next_hint(b, s, t) : BS_NONE;
        if (b.tok && b.c_ntok[t] <= bs_tok_cap(b, s, b.cand[t], next)) {
            // every symbol of this chunk was recorded by the count pass: expand the tokens
            const uint32_t st = expand_tokens_warp(b.tok + bs_tok_base(b, s, b.cand[t]), b.c_ntok[t], cells, b.c_out_off[t],
                                                   &r.out_bytes);
            if (simt::lane() == 0 && b.lb_stats) atomicAdd(&b.lb_stats[3], 1u);
            r.flag = st ? CH_ERR + st : b.c_flag[t];
            if (st) r.out_bytes = b.c_out_len[t];
        } else {
            r = decode_block_chunk<SINK_U16>(sm, b.in_base + b.in_off[s], b.in_size[s], b.cand[t], stop, cells, b.c_out_len[t],
                                             b.c_out_off[t]);
            if (simt::lane() == 0 && b.lb_stats) atomicAdd(&b.lb_stats[4], 1u);
        }
        if (simt::lane() == 0) {
            uint32_t st = ST_OK;
            if (r.flag >= CH_ERR) st = r.flag - CH_ERR;  // e.g. a distance reaching before the stream start
            else if (r.out_bytes != b.c_out_len[t] || r.flag != b.c_flag[t]) st = ST_BAD_CODE;  // cannot happen: same decode twice
            if (st) atomicMax(&b.status[s], st);
        }
        simt::syncwarp();
    }
}

// ---- pieces: a chunk's token run expanded by several warps (lone long streams; cut_token_pieces, bsplit_core.h) ----
// Slot u = region * max_pieces + j. A chunk whose tokens are usable fills up to max_pieces slots of its region with pieces of
// equal token count; a chunk that has to be Huffman-decoded again is one "piece" (p_ntok = BS_PIECE_DECODE); every other slot
// is idle. The resolve kernels then take the slots as their marker domains.
constexpr uint32_t BS_PIECE_DECODE = 0xffffffffu;
constexpr uint32_t BS_PIECE_MIN_TOK = 1024;
struct BsPieces {
    uint32_t max_pieces;
    uint32_t *p_stream;   // per slot
    uint64_t *p_out_off;  // per slot: output offset inside the stream
    uint32_t *p_out_len;  // per slot
    uint32_t *p_flag;     // per slot: CH_IDLE = unused
    uint32_t *p_ntok;     // per slot: tokens of the piece, or BS_PIECE_DECODE
    uint32_t *p_base;     // per stream: first slot
    uint32_t *p_count;    // per stream: slots
};
__device__ __forceinline__ uint32_t bs_piece_tok(uint32_t ntok, uint32_t max_pieces)
{
    uint32_t pt = ((ntok + max_pieces - 1) / max_pieces + 127) & ~127u;
    return pt < BS_PIECE_MIN_TOK ? BS_PIECE_MIN_TOK : pt;
}

__global__ void bs_piece_ranges_kernel(BsBatch b, BsPieces q)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n) return;
    q.p_base[s] = b.chunk_base[s] * q.max_pieces;
    q.p_count[s] = (b.flag[s] && !b.redo[s]) ? b.nchunks[s] * q.max_pieces : 0u;
}

__global__ void __launch_bounds__(BS_WARPS_PER_CTA * 32) bs_pieces_kernel(BsBatch b, BsPieces q)
{
    const uint32_t total_regions = b.summary->total_regions;
    const uint32_t ln = (uint32_t)simt::lane(), M = q.max_pieces;
    const uint32_t warps = gridDim.x * BS_WARPS_PER_CTA;
    for (uint32_t t = blockIdx.x * BS_WARPS_PER_CTA + (threadIdx.x >> 5); t < total_regions; t += warps) {
        const uint32_t s = b.chunk_stream[t], u0 = t * M;
        for (uint32_t j = ln; j < M; j += 32) {
            q.p_stream[u0 + j] = s;
            q.p_out_off[u0 + j] = 0;
            q.p_out_len[u0 + j] = 0;
            q.p_flag[u0 + j] = CH_IDLE;
            q.p_ntok[u0 + j] = 0;
        }
        simt::syncwarp();
        if (!b.flag[s] || b.redo[s] || b.status[s] != ST_OK || b.c_flag[t] == CH_IDLE) continue;
        const uint32_t ntok = b.tok ? b.c_ntok[t] : 0u;
        if (b.tok && ntok && ntok <= bs_tok_cap(b, s, b.cand[t], bs_next_hint(b, s, t))) {
            const uint32_t pt = bs_piece_tok(ntok, M);
            const uint32_t np = cut_token_pieces(b.tok + bs_tok_base(b, s, b.cand[t]), ntok, pt, b.c_out_off[t], q.p_out_off + u0, q.p_out_len + u0);
            for (uint32_t j = ln; j < np; j += 32) {
                q.p_flag[u0 + j] = j + 1 < np ? (uint32_t)CH_RUN : b.c_flag[t];
                q.p_ntok[u0 + j] = j + 1 < np ? pt : ntok - j * pt;
            }
        } else if (ln == 0) {
            q.p_out_off[u0] = b.c_out_off[t];
            q.p_out_len[u0] = b.c_out_len[t];
            q.p_flag[u0] = b.c_flag[t];
            q.p_ntok[u0] = BS_PIECE_DECODE;
        }
        simt::syncwarp();
    }
}

// One warp per slot: tokens -> 16-bit cells of the piece, or the chunk's second Huffman decode.
__global__ void __launch_bounds__(BS_WARPS_PER_CTA * 32) bs_decode_pieces_kernel(BsBatch b, BsPieces q)
{
    const uint32_t M = q.max_pieces, total_slots = b.summary->total_regions * M;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateSmem *sm = reinterpret_cast<InflateSmem *>(smem_raw) + (threadIdx.x >> 5);
    const uint32_t warps = gridDim.x * BS_WARPS_PER_CTA;
    for (uint32_t u = blockIdx.x * BS_WARPS_PER_CTA + (threadIdx.x >> 5); u < total_slots; u += warps) {
        if (q.p_flag[u] == CH_IDLE) continue;
        const uint32_t t = u / M, j = u - t * M, s = b.chunk_stream[t];
        if (b.status[s] != ST_OK && q.p_ntok[u] != BS_PIECE_DECODE) continue;  // (another piece of the stream failed already)
        uint16_t *cells = b.cells + b.cell_base[s] + q.p_out_off[u];
        uint32_t st = ST_OK;
        if (q.p_ntok[u] == BS_PIECE_DECODE) {
            const uint64_t stop = b.c_flag[t] == CH_RUN ? b.exit_bits[t] : BS_NONE;
            const ChunkResult r = decode_block_chunk<SINK_U16>(sm, b.in_base + b.in_off[s], b.in_size[s], b.cand[t], stop, cells,
                                                               b.c_out_len[t], b.c_out_off[t]);
            if (simt::lane() == 0 && b.lb_stats) atomicAdd(&b.lb_stats[4], 1u);
            if (r.flag >= CH_ERR) st = r.flag - CH_ERR;
            else if (r.out_bytes != b.c_out_len[t] || r.flag != b.c_flag[t]) st = ST_BAD_CODE;  // cannot happen: same decode twice
        } else {
            const uint32_t pt = bs_piece_tok(b.c_ntok[t], M);
            uint32_t ob = 0;
            st = expand_tokens_warp(b.tok + bs_tok_base(b, s, b.cand[t]) + (uint64_t)j * pt, q.p_ntok[u], cells, q.p_out_off[u], &ob);
            if (simt::lane() == 0 && b.lb_stats && j == 0) atomicAdd(&b.lb_stats[3], 1u);
            if (!st && ob != q.p_out_len[u]) st = ST_BAD_CODE;  // cannot happen: the cut summed the same lengths
        }
        if (simt::lane() == 0 && st) atomicMax(&b.status[s], st);
        simt::syncwarp();
    }
}

}  // namespace dbg
