// fx_kernels.cuh -- __global__ wrappers of the lane-serial path for single fixed-Huffman-block streams
// (fx_core.h): classify / assign / fill, head (one warp per chunk), sizes (one lane per survivor), chain
// (one warp per stream), tokens (one lane per chunk), expand (one warp per group of chunks). The cells are
// turned into bytes by the resolve kernels of split_kernels.cuh, which see the GROUPS as their chunks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bsplit_core.h"
#include "fx_core.h"
#include "split_kernels.cuh"

namespace dbg {

constexpr uint64_t FX_MIN_BYTES = 8192;   // shorter streams stay on the warp-per-stream path
constexpr int FX_WARPS_PER_CTA = 4;
constexpr int FX_LANE_THREADS = 128;
constexpr uint32_t FX_NONE = 0xffffffffu;

struct FxSummary {            // device -> host after classify and again after chain
    uint32_t n_fx;            // streams on this path
    uint32_t total_chunks;
    uint32_t total_groups;
    uint32_t n_extra;         // survivors beyond the first of their chunk (normally 0)
    uint64_t fx_in;           // compressed bytes of those streams
    uint64_t max_in;          // the longest of them
    uint64_t cells_used;      // exact number of 16-bit cells (chain)
    uint64_t tok_used;        // exact number of tokens (chain)
    uint32_t n_redo;          // streams handed back to the warp-per-stream kernel
    uint32_t pad;
};

struct FxBatch {
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
    uint32_t chunk_bytes;        // compressed bytes per chunk (one lane each)
    uint32_t group_chunks;       // chunks per group (one expansion warp and one marker domain each)
    uint32_t extra_cap;
    uint64_t cells_cap, tok_cap;  // capacities of the cell / token buffers (a stream that does not fit is handed back)
    uint64_t *stats;              // per context, never reset: [0] streams decoded here, [1] handed back, [2] extra survivors
    FxSummary *summary;
    // per stream
    uint32_t *flag;         // 1 = on this path
    uint32_t *redo;         // 1 = handed back (flag stays set)
    uint32_t *chunk_base, *nchunks, *group_base, *ngroups;
    uint64_t *cell_base, *tok_base;
    // per chunk
    uint32_t *chunk_stream;
    uint32_t *hyp;          // x 32
    uint32_t *surv_start;   // x 32
    uint32_t *extra_slot;   // x 32: work item of survivor s >= 1 (survivor 0 of chunk t is item t)
    uint32_t *nsurv;
    uint32_t *c_surv;       // the real survivor, FX_NONE = chunk not reached
    uint32_t *c_out_off, *c_tok_off;
    // work items
    uint32_t *extra_item;   // extra_cap: chunk * 32 + survivor
    FxRec *rec;             // total_chunks + extra_cap
    // per group (the resolve kernels' "chunks")
    uint32_t *group_stream;
    uint64_t *g_out_off;
    uint32_t *g_out_len, *g_flag, *g_ntok;
    uint64_t *g_tok_off;
    // bulk
    uint32_t *tok;
    uint16_t *cells;
};

__global__ void fx_classify_kernel(FxBatch b)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n) return;
    uint32_t flag = 0;
    const uint64_t size = b.in_size[s], cap = b.out_cap[s];
    const bool ok = (!b.pre_status || b.pre_status[s] == 0) && size >= FX_MIN_BYTES && cap >= size && size < (1ull << 31) &&
                    cap < (1ull << 32) - 1024;
    if (ok && is_single_fixed_block(b.in_base + b.in_off[s])) {
        flag = 1;
        atomicAdd(&b.summary->n_fx, 1u);
        atomicAdd((unsigned long long *)&b.summary->fx_in, (unsigned long long)size);
        atomicMax((unsigned long long *)&b.summary->max_in, (unsigned long long)size);
    }
    b.flag[s] = flag;
    b.redo[s] = 0;
}

__global__ void fx_assign_kernel(FxBatch b)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n || !b.flag[s]) return;
    const uint32_t nch = (uint32_t)((b.in_size[s] + b.chunk_bytes - 1) / b.chunk_bytes);
    const uint32_t ng = (nch + b.group_chunks - 1) / b.group_chunks;
    b.chunk_base[s] = atomicAdd(&b.summary->total_chunks, nch);
    b.nchunks[s] = nch;
    b.group_base[s] = atomicAdd(&b.summary->total_groups, ng);
    b.ngroups[s] = ng;
}

__global__ void fx_fill_kernel(FxBatch b)
{
    const uint32_t s = blockIdx.x;
    if (!b.flag[s]) return;
    const uint32_t nch = b.nchunks[s], base = b.chunk_base[s], ng = b.ngroups[s], gb = b.group_base[s];
    for (uint32_t c = threadIdx.x; c < nch; c += blockDim.x) b.chunk_stream[base + c] = s;
    for (uint32_t g = threadIdx.x; g < ng; g += blockDim.x) b.group_stream[gb + g] = s;
}

// Head pass: one warp per chunk.
__global__ void __launch_bounds__(FX_WARPS_PER_CTA * 32) fx_head_kernel(FxBatch b)
{
    __shared__ FxLuts luts;
    fx_build_luts(&luts, threadIdx.x, blockDim.x);
    __syncthreads();
    const uint32_t T = b.summary->total_chunks;
    const uint32_t ln = (uint32_t)simt::lane();
    const uint32_t warps = gridDim.x * FX_WARPS_PER_CTA;
    for (uint32_t t = blockIdx.x * FX_WARPS_PER_CTA + (threadIdx.x >> 5); t < T; t += warps) {
        const uint32_t s = b.chunk_stream[t];
        const uint32_t ns = fx_head_warp(&luts, b.in_base + b.in_off[s], b.in_size[s], t - b.chunk_base[s], b.chunk_bytes,
                                         b.hyp + (uint64_t)t * 32, b.surv_start + (uint64_t)t * 32);
        uint32_t base = 0;
        if (ln == 0) {
            b.nsurv[t] = ns;
            if (ns > 1) {
                base = atomicAdd(&b.summary->n_extra, ns - 1);
                atomicAdd((unsigned long long *)&b.stats[2], (unsigned long long)(ns - 1));
            }
        }
        if (ns > 1) {  // rare: more than one chain is left; every one of them becomes a work item
            base = simt::shfl(base, 0);
            if (ln >= 1 && ln < ns) {
                const uint32_t idx = base + ln - 1;
                if (idx < b.extra_cap) {
                    b.extra_item[idx] = t * 32 + ln;
                    b.extra_slot[(uint64_t)t * 32 + ln] = T + idx;
                } else {
                    b.extra_slot[(uint64_t)t * 32 + ln] = FX_NONE;
                }
            }
        }
        simt::syncwarp();
    }
}

// Sizes pass: one lane per survivor.
__global__ void __launch_bounds__(FX_LANE_THREADS) fx_sizes_kernel(FxBatch b)
{
    __shared__ FxLuts luts;
    fx_build_luts(&luts, threadIdx.x, blockDim.x);
    __syncthreads();
    const uint32_t T = b.summary->total_chunks;
    uint32_t n_extra = b.summary->n_extra;
    if (n_extra > b.extra_cap) n_extra = b.extra_cap;
    const uint32_t total = T + n_extra;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t t = i, sv = 0;
        if (i >= T) {
            const uint32_t it = b.extra_item[i - T];
            t = it >> 5;
            sv = it & 31;
        } else if (b.nsurv[t] == 0) {
            continue;
        }
        const uint32_t s = b.chunk_stream[t], c = t - b.chunk_base[s];
        b.rec[i] = fx_sizes_lane(&luts, b.in_base + b.in_off[s], b.in_size[s], c, b.chunk_bytes, b.surv_start[(uint64_t)t * 32 + sv],
                                 b.hyp + (uint64_t)(t + 1) * 32);
    }
}

// Chain: one warp per stream follows the survivor links from chunk 0. 32 chunks per step; while every
// chunk of a step has exactly one survivor (the normal case) the links are known without walking them.
__global__ void __launch_bounds__(FX_WARPS_PER_CTA * 32) fx_chain_kernel(FxBatch b)
{
    const uint32_t ln = (uint32_t)simt::lane();
    const uint32_t warps = gridDim.x * FX_WARPS_PER_CTA;
    for (uint32_t s = blockIdx.x * FX_WARPS_PER_CTA + (threadIdx.x >> 5); s < b.n; s += warps) {
        if (!b.flag[s]) continue;
        const uint32_t nch = b.nchunks[s], base = b.chunk_base[s];
        uint64_t pos = 0, tok = 0;
        uint32_t entry = 0, end_flag = CH_RUN;
        bool ended = false, redo = false;
        for (uint32_t c0 = 0; c0 < nch; c0 += 32) {
            const uint32_t c = c0 + ln, t = base + c;
            const bool valid = c < nch;
            if (ended || redo) {
                if (valid) b.c_surv[t] = FX_NONE;
                continue;
            }
            const uint32_t ns = valid ? b.nsurv[t] : 1u;
            FxRec r;
            r.out_bytes = r.ntok = r.exit_rel = 0;
            r.link = CH_IDLE;
            uint32_t sv = 0;
            if (!simt::any(ns != 1)) {
                if (valid) r = b.rec[t];
            } else {
                // some chunk of this step has several survivors (or none): walk the links one by one
                uint32_t e = entry;
                for (uint32_t k = 0; k < 32 && c0 + k < nch; k++) {
                    const uint32_t tk = base + c0 + k;
                    const uint32_t item = e == 0 ? (b.nsurv[tk] ? tk : FX_NONE) : b.extra_slot[(uint64_t)tk * 32 + e];
                    if (item == FX_NONE) {  // no room was left for this survivor's work item
                        redo = true;
                        break;
                    }
                    const FxRec rk = b.rec[item];
                    if (k == ln) {
                        r = rk;
                        sv = e;
                    }
                    if ((rk.link & 0xff) != CH_RUN) break;
                    e = rk.link >> 8;
                }
                entry = e;
                if (redo) {
                    if (valid) b.c_surv[t] = FX_NONE;
                    continue;
                }
            }
            const uint32_t fl = r.link & 0xff;
            const uint32_t stops = simt::ballot(valid && fl != CH_RUN && fl != CH_IDLE);
            const uint32_t first = stops ? (uint32_t)simt::ffs(stops) - 1 : 31u;
            const bool active = valid && fl != CH_IDLE && ln <= first;
            uint32_t io = active ? r.out_bytes : 0u, it = active ? r.ntok : 0u;
            const uint32_t mo = io, mt = it;
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t yo = simt::shfl_up(io, d), yt = simt::shfl_up(it, d);
                if (ln >= (uint32_t)d) {
                    io += yo;
                    it += yt;
                }
            }
            if (valid) {
                b.c_surv[t] = active ? sv : FX_NONE;
                b.c_out_off[t] = (uint32_t)(pos + io - mo);
                b.c_tok_off[t] = (uint32_t)(tok + it - mt);
            }
            pos += simt::shfl(io, 31);
            tok += simt::shfl(it, 31);
            if (stops) {
                ended = true;
                end_flag = simt::shfl(fl, (int)first);
            }
        }
        simt::syncwarp();
        uint32_t st = ST_OK;
        if (!ended) st = ST_TRUNCATED;  // cannot happen: the last chunk's run ends at the rule-Q2 limit at the latest
        else if (end_flag >= CH_ERR) st = end_flag - CH_ERR;
        if (st == ST_OK && pos > b.out_cap[s]) st = ST_OUT_OVERFLOW;
        if (st == ST_BAD_CODE) redo = true;  // an internal inconsistency: let the sequential decoder have the last word
        const uint32_t ng = b.ngroups[s], gb = b.group_base[s], G = b.group_chunks;
        for (uint32_t g = ln; g < ng; g += 32) {
            const uint32_t tf = base + g * G;
            uint32_t len = 0, nt = 0, gf = CH_IDLE;
            uint64_t go = 0, gt = 0;
            if (st == ST_OK && !redo && b.c_surv[tf] != FX_NONE) {
                go = b.c_out_off[tf];
                gt = b.c_tok_off[tf];
                const bool more = (g + 1) * G < nch && b.c_surv[tf + G] != FX_NONE;
                len = (uint32_t)((more ? b.c_out_off[tf + G] : pos) - go);
                nt = (uint32_t)((more ? b.c_tok_off[tf + G] : tok) - gt);
                gf = more ? (uint32_t)CH_RUN : end_flag;
            }
            b.g_out_off[gb + g] = go;
            b.g_tok_off[gb + g] = gt;
            b.g_out_len[gb + g] = len;
            b.g_ntok[gb + g] = nt;
            b.g_flag[gb + g] = gf;
        }
        if (ln == 0) {
            uint64_t cb = 0, tb = 0;
            if (!redo && st == ST_OK) {
                // room in the cell / token buffers? (sized exactly by the host when it read the counters back, by an
                // upper bound / an estimate when it did not wait for them)
                cb = atomicAdd((unsigned long long *)&b.summary->cells_used, (unsigned long long)pos);
                tb = atomicAdd((unsigned long long *)&b.summary->tok_used, (unsigned long long)tok);
                if (cb + pos > b.cells_cap || tb + tok > b.tok_cap) redo = true;
            }
            if (redo) {
                b.redo[s] = 1;
                b.cell_base[s] = 0;
                b.tok_base[s] = 0;
                atomicAdd(&b.summary->n_redo, 1u);
                atomicAdd((unsigned long long *)&b.stats[1], 1ull);
            } else {
                b.status[s] = st;
                b.out_size[s] = st == ST_OK ? pos : 0;
                b.cell_base[s] = cb;
                b.tok_base[s] = tb;
                atomicAdd((unsigned long long *)&b.stats[0], 1ull);
            }
        }
        redo = simt::shfl(redo ? 1u : 0u, 0) != 0;
        if (redo) {  // (decided after the groups were written) nothing of this stream may be expanded
            for (uint32_t g = ln; g < ng; g += 32) {
                b.g_flag[gb + g] = CH_IDLE;
                b.g_out_len[gb + g] = 0;
                b.g_ntok[gb + g] = 0;
            }
        }
        simt::syncwarp();
    }
}

// Token pass: one lane per chunk, the real survivor only.
__global__ void __launch_bounds__(FX_LANE_THREADS) fx_tokens_kernel(FxBatch b)
{
    __shared__ FxLuts luts;
    fx_build_luts(&luts, threadIdx.x, blockDim.x);
    __syncthreads();
    const uint32_t T = b.summary->total_chunks;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        const uint32_t sv = b.c_surv[t];
        if (sv == FX_NONE) continue;
        const uint32_t s = b.chunk_stream[t];
        if (b.redo[s] || b.status[s] != ST_OK) continue;
        const uint32_t item = sv == 0 ? t : b.extra_slot[(uint64_t)t * 32 + sv];
        const FxRec r = b.rec[item];
        uint32_t ob = 0, nt = 0;
        fx_tokens_lane(&luts, b.in_base + b.in_off[s], b.in_size[s], t - b.chunk_base[s], b.chunk_bytes,
                       b.surv_start[(uint64_t)t * 32 + sv], r.exit_rel, b.tok + b.tok_base[s] + b.c_tok_off[t], &ob, &nt);
        if (ob != r.out_bytes || nt != r.ntok) atomicMax(&b.status[s], (uint32_t)ST_BAD_CODE);  // cannot happen: same run twice
    }
}

// Expansion: one warp per group, tokens -> 16-bit cells.
__global__ void __launch_bounds__(FX_WARPS_PER_CTA * 32) fx_expand_kernel(FxBatch b)
{
    const uint32_t NG = b.summary->total_groups;
    const uint32_t warps = gridDim.x * FX_WARPS_PER_CTA;
    for (uint32_t g = blockIdx.x * FX_WARPS_PER_CTA + (threadIdx.x >> 5); g < NG; g += warps) {
        if (b.g_flag[g] == CH_IDLE) continue;
        const uint32_t s = b.group_stream[g];
        if (b.redo[s] || b.status[s] != ST_OK) continue;
        uint32_t ob = 0;
        const uint32_t st = expand_tokens_warp(b.tok + b.tok_base[s] + b.g_tok_off[g], b.g_ntok[g],
                                               b.cells + b.cell_base[s] + b.g_out_off[g], b.g_out_off[g], &ob);
        if (simt::lane() == 0) {
            if (st) atomicMax(&b.status[s], st);  // a distance that reaches before the start of the stream
            else if (ob != b.g_out_len[g]) atomicMax(&b.status[s], (uint32_t)ST_BAD_CODE);  // cannot happen
        }
        simt::syncwarp();
    }
}

}  // namespace dbg
