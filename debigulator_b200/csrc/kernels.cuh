// kernels.cuh -- __global__ entry points of the decode path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "inflate_core.h"

namespace dbg {

constexpr int INFLATE_WARPS_PER_CTA = 4;
constexpr int INFLATE_THREADS = INFLATE_WARPS_PER_CTA * 32;
#ifndef DBG_BUILD_INFLATE_CTAS
#define DBG_BUILD_INFLATE_CTAS 6
#endif
constexpr int INFLATE_CTAS_PER_SM = DBG_BUILD_INFLATE_CTAS;  // resident CTAs per SM = the register budget (80 at 6). Round 1 (symbol walk only): 6 beat 8 (L2 footprint of the 32 KiB windows). Round 2 (lane-parallel rounds): cfg2 4 -> 50.2 ms, 5 -> 37.2 ms, 6 -> 38.8 ms, but the config-5 shape 5 -> 180 ms, 6 -> 112 ms (the block-split kernels size their grids and their threshold by this number)

struct InflateBatch {
    const uint8_t *in_base;
    const uint64_t *in_off;
    const uint64_t *in_size;
    uint8_t *out_base;
    const uint64_t *out_off;
    const uint64_t *out_cap;
    uint64_t *out_size;
    uint32_t *status;
    const uint32_t *pre_status;  // optional: non-zero entries are reported as-is and skipped
    const uint32_t *skip;        // optional: non-zero entries are handled elsewhere (split-stream path)
    const uint32_t *skip2;       // optional: same, block-split path
    const uint32_t *only;        // optional: only entries with a non-zero flag are processed (second pass)
    const uint32_t *order;       // optional scheduling permutation
    uint32_t *counter;           // work-queue head, zeroed before launch
    uint32_t n;
    uint32_t *tok_scratch;       // optional: LB_ROUND_TOKENS slots per warp of the grid (lane-parallel rounds, inflate_core.h)
    uint32_t *lb_stats;          // optional counters of those rounds
    uint32_t round_bits;         // their length in bits
};

// Persistent warps pulling streams from a global queue: one warp decodes one
// stream start to finish (inflate_core.h), then fetches the next index.
__global__ void __launch_bounds__(INFLATE_THREADS, INFLATE_CTAS_PER_SM) inflate_batch_kernel(InflateBatch a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateSmem *sm = reinterpret_cast<InflateSmem *>(smem_raw) + (threadIdx.x >> 5);
    const int ln = simt::lane();
    uint32_t *scratch = a.tok_scratch ? a.tok_scratch + (uint64_t)(blockIdx.x * INFLATE_WARPS_PER_CTA + (threadIdx.x >> 5)) * LB_ROUND_TOKENS : nullptr;
    for (;;) {
        uint32_t idx = 0;
        if (ln == 0) idx = atomicAdd(a.counter, 1u);
        idx = simt::shfl(idx, 0);
        if (idx >= a.n) break;
        const uint32_t s = a.order ? a.order[idx] : idx;
        if ((a.skip && a.skip[s]) || (a.skip2 && a.skip2[s]) || (a.only && !a.only[s])) continue;
        uint32_t st = a.pre_status ? a.pre_status[s] : 0u;
        uint64_t fs = 0;
        if (st == 0)
            st = inflate_warp(sm, a.in_base + a.in_off[s], a.in_size[s], a.out_base + a.out_off[s], a.out_cap[s], &fs, scratch, a.lb_stats, a.round_bits);
        if (ln == 0) {
            a.out_size[s] = fs;
            a.status[s] = st;
        }
        simt::syncwarp();
    }
}

// Largest-first work-queue order for callers that do not bring one: a counting sort of the streams by
// the quarter-octave of their estimated decode time (one CTA; order inside a bucket is arbitrary).
// Estimate: compressed bytes (1/64 of them when the stream opens with a stored block, i.e. a plain
// copy) + output capacity / 32 (match copying of highly compressible streams).
constexpr int SCHED_BUCKETS = 256;
__global__ void __launch_bounds__(1024) sched_order_kernel(InflateBatch a, uint32_t *order)
{
    __shared__ uint32_t hist[SCHED_BUCKETS];
    for (uint32_t i = threadIdx.x; i < SCHED_BUCKETS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    auto bucket = [&](uint32_t s) -> uint32_t {
        const uint64_t size = a.in_size[s];
        uint64_t w = size;
        if (size && ((a.in_base[a.in_off[s]] >> 1) & 3) == 0) w = size >> 6;
        w += a.out_cap[s] >> 5;
        if (w < 4) return SCHED_BUCKETS - 1;
        const int e = 63 - __clzll((long long)w);                  // 2 .. 63
        const uint32_t q = (uint32_t)(w >> (e - 2)) & 3;           // two bits below the leading one
        const uint32_t b = (uint32_t)e * 4 + q;                    // larger = heavier
        return b >= SCHED_BUCKETS ? 0 : SCHED_BUCKETS - 1 - b;     // bucket 0 = heaviest
    };
    for (uint32_t s = threadIdx.x; s < a.n; s += blockDim.x) atomicAdd(&hist[bucket(s)], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < SCHED_BUCKETS; i++) {
            const uint32_t c = hist[i];
            hist[i] = run;
            run += c;
        }
    }
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < a.n; s += blockDim.x) order[atomicAdd(&hist[bucket(s)], 1u)] = s;
}

// gzip member framing, one thread per member: the header walk of
// decode_gz.c:123-233 (silent build: FNAME skipped, FCOMMENT not) and the
// payload size rule of decode_gz.c:270 (everything but the 8-byte trailer).
__global__ void gz_scan_kernel(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_size, uint32_t n,
                               uint64_t *payload_off, uint64_t *payload_size, uint32_t *pre_status)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = in_base + in_off[i];
    uint64_t size = in_size[i];
    payload_off[i] = in_off[i];
    payload_size[i] = 0;
    uint32_t st = ST_OK;
    if (size < 10 || p[0] != 31 || p[1] != 139 || p[2] != 8) {
        st = ST_CONTAINER;
    } else {
        uint64_t at = 10, left = size - 10;
        if ((p[3] >> 3) & 1) {
            uint64_t k = 0;
            while (k < left && p[at + k] != 0) k++;
            if (k + 1 > left) {
                st = ST_CONTAINER;  // unterminated FNAME: the reference's size underflows and inflate rejects it
            } else {
                at += k + 1;
                left -= k + 1;
            }
        }
        if (st == ST_OK) {
            if (left < 8) {
                st = ST_CONTAINER;
            } else {
                payload_off[i] = in_off[i] + at;
                payload_size[i] = left - 8;
            }
        }
    }
    pre_status[i] = st;
}

}  // namespace dbg
