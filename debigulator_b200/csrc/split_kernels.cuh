// split_kernels.cuh -- the resolve kernels shared by the chunk-parallel paths (fx_kernels.cuh: single fixed-Huffman
// blocks, one lane per chunk; bsplit_kernels.cuh: long multi-block streams): 16-bit cells -> bytes. Their "chunks" are
// the marker domains of those paths (groups of lane chunks / stretches between block boundaries).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "inflate_core.h"

namespace dbg {

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
    // scratch
    uint32_t *split_flag;   // per stream: 1 = split path
    const uint32_t *redo;   // optional (block-split path): 1 = handed back, nothing to resolve
    uint32_t *chunk_base;   // per stream: first global chunk index
    uint32_t *nchunks;      // per stream: number of chunks
    uint64_t *cell_base;    // per stream: first cell
    uint32_t *chunk_stream; // per chunk
    uint64_t *c_out_off;    // per chunk: output offset inside the stream
    uint32_t *c_out_len;    // per chunk
    uint32_t *c_flag;       // per chunk
    uint16_t *cells;
};


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
