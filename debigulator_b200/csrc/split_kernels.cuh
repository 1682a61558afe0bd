// split_kernels.cuh -- __global__ wrappers of the split-stream path (see
// "split stream" in inflate_core.h): chunk-parallel decode of streams that are a
// single fixed-Huffman block, used when a batch has too few streams to fill the
// GPU with one warp per stream (BASELINE config 4: 32 images of 256 MiB per GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "inflate_core.h"

namespace dbg {

constexpr uint64_t SPLIT_MIN_BYTES = 4 * CHUNK_BYTES;  // smaller streams stay on the warp-per-stream path
constexpr int SPLIT_WARPS_PER_CTA = 4;

struct SplitSummary {      // written by split_classify_kernel, read back by the host
    uint32_t n_split;      // streams taking the split path
    uint32_t total_chunks;
    uint64_t cells_cap;    // upper bound of 16-bit cells needed (sum of output capacities)
    uint64_t cells_used;   // running allocation cursor (split_chain_kernel)
    uint64_t split_in;     // compressed bytes of the streams taking the split path
    uint64_t max_in;       // the longest of them
};

struct SplitBatch {
    const uint8_t *in_base;
    const uint64_t *in_off;
    const uint64_t *in_size;
    uint8_t *out_base;
    const uint64_t *out_off;
    const uint64_t *out_cap;
    uint64_t *out_size;
    uint32_t *status;
    const uint32_t *pre_status;  // optional
    uint32_t n;
    uint32_t chunk_bytes;   // compressed bytes per chunk (CHUNK_BYTES << k)
    // scratch
    SplitSummary *summary;
    uint32_t *split_flag;   // per stream: 1 = split path
    const uint32_t *redo;   // optional (block-split path): 1 = handed back, nothing to resolve
    uint32_t *chunk_base;   // per stream: first global chunk index
    uint32_t *nchunks;      // per stream: number of chunks
    uint64_t *cell_base;    // per stream: first cell
    uint32_t *chunk_stream; // per chunk
    uint64_t *entry_bits;   // per chunk: exact entry (stream-relative bit)
    uint64_t *c_out_off;    // per chunk: output offset inside the stream
    uint32_t *c_out_len;    // per chunk
    uint32_t *c_flag;       // per chunk
    TransferEntry *tf;      // per chunk x 32
    uint16_t *cells;
};


// Which streams take the split path, and how many chunks / cells that needs.
__global__ void split_classify_kernel(SplitBatch b)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n) return;
    uint32_t flag = 0;
    uint64_t size = b.in_size[s], cap = b.out_cap[s];
    bool ok = (!b.pre_status || b.pre_status[s] == 0) && size >= SPLIT_MIN_BYTES && cap >= size && size < (1ull << 31) &&
              cap < (1ull << 32) - 1024;
    if (ok && is_single_fixed_block(b.in_base + b.in_off[s])) {
        flag = 1;
        atomicAdd(&b.summary->n_split, 1u);
        atomicAdd((unsigned long long *)&b.summary->cells_cap, (unsigned long long)cap);
        atomicAdd((unsigned long long *)&b.summary->split_in, (unsigned long long)size);
        atomicMax((unsigned long long *)&b.summary->max_in, (unsigned long long)size);
    }
    b.split_flag[s] = flag;
}

// Chunks per stream, once the host has chosen the chunk size.
__global__ void split_assign_kernel(SplitBatch b)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n || !b.split_flag[s]) return;
    const uint32_t nch = (uint32_t)((b.in_size[s] + b.chunk_bytes - 1) / b.chunk_bytes);
    b.chunk_base[s] = atomicAdd(&b.summary->total_chunks, nch);
    b.nchunks[s] = nch;
}

__global__ void split_fill_kernel(SplitBatch b)
{
    uint32_t s = blockIdx.x;
    if (!b.split_flag[s]) return;
    uint32_t nch = b.nchunks[s], base = b.chunk_base[s];
    for (uint32_t c = threadIdx.x; c < nch; c += blockDim.x) b.chunk_stream[base + c] = s;
}

// Transfer tables: one warp per chunk, one lane per entry-offset hypothesis.
__global__ void __launch_bounds__(SPLIT_WARPS_PER_CTA * 32) split_transfer_kernel(SplitBatch b)
{
    const uint32_t total_chunks = b.summary->total_chunks;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateSmem *sm = reinterpret_cast<InflateSmem *>(smem_raw) + (threadIdx.x >> 5);
    const uint32_t warps = gridDim.x * SPLIT_WARPS_PER_CTA;
    for (uint32_t t = blockIdx.x * SPLIT_WARPS_PER_CTA + (threadIdx.x >> 5); t < total_chunks; t += warps) {
        const uint32_t s = b.chunk_stream[t];
        transfer_chunk_warp(sm, b.in_base + b.in_off[s], b.in_size[s], t - b.chunk_base[s], b.chunk_bytes, b.tf + (uint64_t)t * 32);
        simt::syncwarp();
    }
}

// Exact entry and output offset of every chunk: one thread per stream.
__global__ void split_chain_kernel(SplitBatch b)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n || !b.split_flag[s]) return;
    const uint32_t nch = b.nchunks[s], base = b.chunk_base[s];
    uint64_t pos = 0;
    uint32_t idx = 0, st = ST_OK;
    bool ended = false;
    for (uint32_t c = 0; c < nch; c++) {
        const uint32_t t = base + c;
        if (ended) {
            b.c_flag[t] = CH_IDLE;
            b.c_out_len[t] = 0;
            b.c_out_off[t] = pos;
            b.entry_bits[t] = 0;
            continue;
        }
        const TransferEntry e = b.tf[(uint64_t)t * 32 + idx];
        b.entry_bits[t] = c == 0 ? 3 : (uint64_t)c * b.chunk_bytes * 8 + idx;
        b.c_out_off[t] = pos;
        b.c_out_len[t] = e.out_bytes;
        b.c_flag[t] = e.flag;
        pos += e.out_bytes;
        if (e.flag != CH_RUN) {
            ended = true;
            if (e.flag >= CH_ERR) st = ST_BAD_SYMBOL;
        }
        idx = e.next;
    }
    if (!ended) st = ST_TRUNCATED;
    if (st == ST_OK && pos > b.out_cap[s]) st = ST_OUT_OVERFLOW;
    b.status[s] = st;
    b.out_size[s] = st == ST_OK ? pos : 0;
    b.cell_base[s] = st == ST_OK ? atomicAdd((unsigned long long *)&b.summary->cells_used, (unsigned long long)pos) : 0;
}

// Chunk decode into 16-bit cells: one warp per chunk.
__global__ void __launch_bounds__(SPLIT_WARPS_PER_CTA * 32) split_decode_kernel(SplitBatch b)
{
    const uint32_t total_chunks = b.summary->total_chunks;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateSmem *sm = reinterpret_cast<InflateSmem *>(smem_raw) + (threadIdx.x >> 5);
    const uint32_t warps = gridDim.x * SPLIT_WARPS_PER_CTA;
    for (uint32_t t = blockIdx.x * SPLIT_WARPS_PER_CTA + (threadIdx.x >> 5); t < total_chunks; t += warps) {
        const uint32_t s = b.chunk_stream[t];
        if (b.status[s] != ST_OK || b.c_flag[t] == CH_IDLE) continue;
        const uint32_t c = t - b.chunk_base[s];
        ChunkResult r = decode_chunk<SINK_U16>(sm, b.in_base + b.in_off[s], b.in_size[s], c, b.chunk_bytes, b.entry_bits[t],
                                               b.cells + b.cell_base[s] + b.c_out_off[t], b.c_out_len[t], b.c_out_off[t]);
        if (simt::lane() == 0) {
            uint32_t st = ST_OK;
            if (r.flag >= CH_ERR) st = r.flag - CH_ERR;
            else if (r.out_bytes != b.c_out_len[t] || r.flag != b.c_flag[t]) st = ST_BAD_CODE;  // cannot happen: tables are exact
            if (st) atomicMax(&b.status[s], st);
        }
        simt::syncwarp();
    }
}

// Cells -> bytes. A marker in chunk c points into the 32 KiB of output that precede the chunk,
// i.e. into the TAIL (last 32 KiB) of the chunks before it. Only the tails therefore form a serial
// chain; they are resolved chunk after chunk by one CTA per stream (32 cells per thread, the next
// chunk's cells are already in flight while the current one is resolved). Everything else -- the
// body of every chunk -- is then resolved by all CTAs at once against the finished tails.
constexpr int RESOLVE_THREADS = 1024;
constexpr uint32_t TAIL_BYTES = 32768;
constexpr int TAIL_PER_THREAD = TAIL_BYTES / RESOLVE_THREADS;

__global__ void __launch_bounds__(RESOLVE_THREADS) split_resolve_tails_kernel(SplitBatch b)
{
    const uint32_t s = blockIdx.x;
    if (b.redo && b.redo[s]) return;
    if (!b.split_flag[s] || b.status[s] != ST_OK) {
        if (b.split_flag[s] && threadIdx.x == 0) b.out_size[s] = 0;
        return;
    }
    const uint32_t nch = b.nchunks[s], base = b.chunk_base[s];
    uint8_t *out = b.out_base + b.out_off[s];
    const uint16_t *cells = b.cells + b.cell_base[s];
    uint32_t v[TAIL_PER_THREAD / 2], nv[TAIL_PER_THREAD / 2];  // two 16-bit cells per register
    uint64_t o = 0, t0 = 0;
    uint32_t tl = 0;
    auto fetch = [&](uint32_t c, uint32_t *dst, uint64_t &co, uint64_t &cs, uint32_t &cl) {
        co = b.c_out_off[base + c];
        const uint32_t len = b.c_out_len[base + c];
        cl = len < TAIL_BYTES ? len : TAIL_BYTES;
        cs = co + len - cl;
#pragma unroll
        for (int k = 0; k < TAIL_PER_THREAD / 2; k++) {
            uint32_t i0 = threadIdx.x + (2 * k) * RESOLVE_THREADS, i1 = i0 + RESOLVE_THREADS;
            uint32_t lo = i0 < cl ? cells[cs + i0] : 0, hi = i1 < cl ? cells[cs + i1] : 0;
            dst[k] = lo | (hi << 16);
        }
    };
    fetch(0, v, o, t0, tl);
    for (uint32_t c = 0; c < nch; c++) {
        uint64_t no = 0, nt0 = 0;
        uint32_t ntl = 0;
        if (c + 1 < nch) fetch(c + 1, nv, no, nt0, ntl);
        // all marker loads first, then all stores: a marker only points before this chunk, so nothing read here
        // is written here, but the compiler cannot know that and would otherwise chain load -> store -> load,
        // one L2 round trip per cell (measured: 10.6 us per chunk, 49 ms for four 150 MB streams)
#pragma unroll
        for (int k = 0; k < TAIL_PER_THREAD; k++) {
            const uint32_t i = threadIdx.x + k * RESOLVE_THREADS;
            const uint32_t x = (v[k >> 1] >> (16 * (k & 1))) & 0xffff;
            if (i < tl && x >= 256) {  // marker: 256 + 32768 + (negative source index)
                const uint32_t y = out[o + x - 33024];
                v[k >> 1] = (v[k >> 1] & ~(0xffffu << (16 * (k & 1)))) | (y << (16 * (k & 1)));
            }
        }
#pragma unroll
        for (int k = 0; k < TAIL_PER_THREAD; k++) {
            const uint32_t i = threadIdx.x + k * RESOLVE_THREADS;
            if (i < tl) out[t0 + i] = (uint8_t)(v[k >> 1] >> (16 * (k & 1)));
        }
        __syncthreads();  // the next chunk's markers may point at these bytes
#pragma unroll
        for (int k = 0; k < TAIL_PER_THREAD / 2; k++) v[k] = nv[k];
        o = no;
        t0 = nt0;
        tl = ntl;
    }
}

__global__ void __launch_bounds__(256) split_resolve_body_kernel(SplitBatch b, uint32_t total_chunks)
{
    for (uint32_t t = blockIdx.x; t < total_chunks; t += gridDim.x) {
        const uint32_t s = b.chunk_stream[t];
        if (b.status[s] != ST_OK) continue;
        const uint32_t len = b.c_out_len[t];
        if (len <= TAIL_BYTES) continue;
        const uint64_t o = b.c_out_off[t];
        const uint32_t body = len - TAIL_BYTES;
        uint8_t *out = b.out_base + b.out_off[s];
        const uint16_t *cells = b.cells + b.cell_base[s];
        // four cells per thread step, loads before stores (a marker points into the finished tails, never
        // into a body, so the order is free)
        for (uint32_t i0 = threadIdx.x; i0 < body; i0 += 4 * 256) {
            uint32_t x[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = i0 + k * 256;
                x[k] = i < body ? cells[o + i] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x[k] >= 256) x[k] = out[o + x[k] - 33024];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t i = i0 + k * 256;
                if (i < body) out[o + i] = (uint8_t)x[k];
            }
        }
    }
}

}  // namespace dbg
