// dbg_api.cu -- host side of libdebigulator_b200.so: context, batched C-ABI
// (include/debigulator_b200.h) and kernel launches. No CPU decode path exists
// in this library: every entry point needs a CUDA device and reports
// DBG_ERR_NO_DEVICE / DBG_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <utility>
#include <vector>

#include "../../include/debigulator_b200.h"
#include "bmp_kernels.cuh"
#include "bsplit_kernels.cuh"
#include "fx_kernels.cuh"
#include "kernels.cuh"
#include "png_kernels.cuh"
#include "split_kernels.cuh"

static thread_local char g_err[512] = "";

static void set_err(dbg_ctx *ctx, const char *fmt, ...);

// Grow-only buffer (device or pinned host).
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    bool pinned_host = false;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        release();
        size_t want = n + n / 8 + 256;
        cudaError_t e = pinned_host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            cap = 0;
            return e;
        }
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (p) {
            if (pinned_host) cudaFreeHost(p);
            else cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
};

struct dbg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    static constexpr int MAX_WAVES = 16;
    cudaStream_t wave_stream[MAX_WAVES] = {};
    int waves = 16;  // waves the packed host API cuts a large batch into (H2D / kernels / D2H overlap); cfg2 end to end: 2 -> 23.5, 4 -> 25.8, 8 -> 26.7, 16 -> 27.4 GB/s
    cudaEvent_t wave_ready = nullptr;
    cudaStream_t aux_stream = nullptr;  // the warp-per-stream kernel runs here beside the block-split kernels
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    uint64_t launches = 0;
    char err[512] = "";
    // optional per-launch timing of the dominant (inflate) kernel, for roofline reports
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    // device-side scratch
    Buf d_counter;                 // work-queue heads
    Buf d_meta;                    // derived descriptors (gzip payloads, PNG streams)
    Buf d_png_scratch;             // compacted IDAT + filtered scanlines
    Buf d_split, d_split_chunks, d_cells, h_summary;  // split-stream path: chunk tables, 16-bit cells, pinned summary
    // block-split path (long multi-block streams): per wave slot, since waves run concurrently
    Buf d_sched[MAX_WAVES];        // work-queue order computed on the device when the caller brings none
    uint32_t split_chunk_forced = 0;  // DBG_SPLIT_CHUNK: fixed chunk size of the split-stream path (experiments)
    bool bsplit_allowed = true;    // cleared while the packed API runs several waves at once
    Buf d_bs_stream[MAX_WAVES], d_bs_region[MAX_WAVES], d_bs_cells[MAX_WAVES], d_bs_tok[MAX_WAVES], h_bs_summary[MAX_WAVES];
    uint32_t bsplit_tok_per_byte = 4;             // token slots per compressed byte (0 = no tokens: decode twice)
    uint64_t bsplit_tok_max_bytes = 24ull << 30;  // the token area never grows beyond this
    bool bsplit = true;
    uint64_t bsplit_min_bytes = dbg::BS_MIN_BYTES;
    uint32_t bsplit_factor_q = 8;
    uint32_t bsplit_region = dbg::REGION_BYTES, bsplit_region_min = 16384;
    uint64_t bs_streams = 0, bs_fallbacks = 0;  // counters: streams that took the block-split path / were handed back
    uint32_t split_max_streams = 1536;  // packed host API: smaller batches run as one wave (so that the per-context paths may take them)
    bool fx = true;                     // lane-serial path for single fixed-Huffman-block streams (fx_kernels.cuh)
    uint32_t fx_group_forced = 0;       // DBG_FX_GROUP: fixed group size (experiments)
    uint64_t fx_streams = 0, fx_redo = 0, fx_extra = 0;  // counters: streams on that path / handed back / extra survivors
    Buf d_fx_tok;                       // its tokens
    bool verify = false;                // opt-in: check gzip CRC32 / ISIZE trailers
    uint32_t inflate_ctas_per_sm = dbg::INFLATE_CTAS_PER_SM;  // resident streams per SM = 4x this (tunable: L2 footprint)
    // host-API staging
    Buf d_in, d_out, d_desc;       // arenas + descriptor tables
    Buf h_in, h_out, h_desc;       // pinned mirrors
    dbg_ctx()
    {
        h_in.pinned_host = h_out.pinned_host = h_desc.pinned_host = h_summary.pinned_host = true;
        for (int k = 0; k < MAX_WAVES; k++) h_bs_summary[k].pinned_host = true;
    }
};

static void set_err(dbg_ctx *ctx, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (ctx) memcpy(ctx->err, g_err, sizeof(g_err));
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            set_err(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return DBG_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

extern "C" int dbg_version(void) { return 100; }

extern "C" int dbg_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_err(nullptr, "cudaGetDeviceCount: %s -- this library has no CPU path", cudaGetErrorString(e));
        return DBG_ERR_NO_DEVICE;
    }
    return n;
}

extern "C" const char *dbg_last_error(const dbg_ctx *ctx) { return ctx ? ctx->err : g_err; }
extern "C" int dbg_ctx_device(const dbg_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" uint64_t dbg_kernel_launches(const dbg_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" dbg_ctx *dbg_create(int device)
{
    int n = dbg_device_count();
    if (n <= 0) {
        if (n == 0) set_err(nullptr, "no CUDA device visible -- this library has no CPU path");
        return nullptr;
    }
    if (device < 0 || device >= n) {
        set_err(nullptr, "device %d out of range (0..%d)", device, n - 1);
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_err(nullptr, "cannot select device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    if (prop.major < 10) {
        set_err(nullptr, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return nullptr;
    }
    dbg_ctx *ctx = new dbg_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_err(nullptr, "cudaStreamCreate: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return nullptr;
    }
    for (int i = 0; i < dbg_ctx::MAX_WAVES; i++) cudaStreamCreateWithFlags(&ctx->wave_stream[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->wave_ready, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming);
    size_t smem = sizeof(dbg::InflateSmem) * dbg::INFLATE_WARPS_PER_CTA;
    cudaFuncSetAttribute(dbg::inflate_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(dbg::inflate_batch_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 85);
    dbg::png_configure_kernels();
    if (const char *e = getenv("DBG_SPLIT_MAX_STREAMS")) ctx->split_max_streams = (uint32_t)atoi(e);
    if (const char *e = getenv("DBG_SPLIT_CHUNK")) ctx->split_chunk_forced = (uint32_t)std::min(1 << 20, std::max((int)dbg::FX_MIN_CHUNK, atoi(e))) & ~15u;
    if (const char *e = getenv("DBG_FX_GROUP")) ctx->fx_group_forced = (uint32_t)std::min(1 << 22, std::max(4096, atoi(e))) & ~15u;
    if (const char *e = getenv("DBG_FX")) ctx->fx = atoi(e) != 0;
    if (const char *e = getenv("DBG_WAVES")) ctx->waves = std::min((int)dbg_ctx::MAX_WAVES, std::max(1, atoi(e)));
    if (const char *e = getenv("DBG_BSPLIT")) ctx->bsplit = atoi(e) != 0;
    if (const char *e = getenv("DBG_BSPLIT_FACTOR_Q")) ctx->bsplit_factor_q = (uint32_t)std::max(1, atoi(e));
    if (const char *e = getenv("DBG_BSPLIT_TOKENS")) ctx->bsplit_tok_per_byte = (uint32_t)std::min(8, std::max(0, atoi(e)));
    if (const char *e = getenv("DBG_BSPLIT_REGION")) ctx->bsplit_region = (uint32_t)std::min(1 << 20, std::max(4096, atoi(e)));
    if (const char *e = getenv("DBG_BSPLIT_REGION_MIN")) ctx->bsplit_region_min = (uint32_t)std::min((int)ctx->bsplit_region, std::max(4096, atoi(e)));
    if (const char *e = getenv("DBG_BSPLIT_MIN_BYTES")) ctx->bsplit_min_bytes = std::max<uint64_t>(strtoull(e, nullptr, 10), 2 * (uint64_t)ctx->bsplit_region);
    if (const char *e = getenv("DBG_INFLATE_CTAS_PER_SM")) {
        int v = atoi(e);
        if (v >= 1 && v <= dbg::INFLATE_CTAS_PER_SM) ctx->inflate_ctas_per_sm = (uint32_t)v;
    }
    return ctx;
}

extern "C" void dbg_destroy(dbg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    Buf *all[] = {&ctx->d_counter, &ctx->d_meta, &ctx->d_png_scratch, &ctx->d_in,    &ctx->d_out,    &ctx->d_desc,
                  &ctx->h_in,      &ctx->h_out,  &ctx->h_desc,        &ctx->d_split, &ctx->d_cells, &ctx->h_summary,
                  &ctx->d_split_chunks, &ctx->d_fx_tok};
    for (Buf *b : all) b->release();
    for (int i = 0; i < dbg_ctx::MAX_WAVES; i++) {
        ctx->d_sched[i].release();
        ctx->d_bs_stream[i].release();
        ctx->d_bs_region[i].release();
        ctx->d_bs_cells[i].release();
        ctx->d_bs_tok[i].release();
        ctx->h_bs_summary[i].release();
        if (ctx->wave_stream[i]) cudaStreamDestroy(ctx->wave_stream[i]);
    }
    if (ctx->wave_ready) cudaEventDestroy(ctx->wave_ready);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
    if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int dbg_bsplit_stats(const dbg_ctx *ctx, uint64_t *streams, uint64_t *fallbacks)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (streams) *streams = ctx->bs_streams;
    if (fallbacks) *fallbacks = ctx->bs_fallbacks;
    return DBG_OK;
}

extern "C" int dbg_fx_stats(const dbg_ctx *ctx, uint64_t *streams, uint64_t *handed_back, uint64_t *extra_runs)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (streams) *streams = ctx->fx_streams;
    if (handed_back) *handed_back = ctx->fx_redo;
    if (extra_runs) *extra_runs = ctx->fx_extra;
    return DBG_OK;
}

extern "C" int dbg_set_verify(dbg_ctx *ctx, int on)
{
    if (!ctx) return DBG_ERR_ARG;
    ctx->verify = on != 0;
    return DBG_OK;
}

extern "C" int dbg_profile_enable(dbg_ctx *ctx, int on)
{
    if (!ctx) return DBG_ERR_ARG;
    ctx->profiling = on != 0;
    ctx->prof_used = 0;
    return DBG_OK;
}

// Sum of the device durations (ms) of the inflate kernels launched since
// dbg_profile_enable(ctx, 1), and how many there were. Waits for them.
extern "C" int dbg_profile_read(dbg_ctx *ctx, double *total_ms, uint64_t *launches)
{
    if (!ctx || !total_ms || !launches) return DBG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    double sum = 0;
    for (size_t i = 0; i < ctx->prof_used; i++) {
        float ms = 0;
        CU(cudaEventSynchronize(ctx->prof_events[i].second));
        CU(cudaEventElapsedTime(&ms, ctx->prof_events[i].first, ctx->prof_events[i].second));
        sum += ms;
    }
    *total_ms = sum;
    *launches = ctx->prof_used;
    ctx->prof_used = 0;
    return DBG_OK;
}

extern "C" int dbg_synchronize(dbg_ctx *ctx)
{
    if (!ctx) return DBG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return DBG_OK;
}

// ------------------------------------------------------------------ launches --
static int launch_inflate_plain(dbg_ctx *ctx, dbg::InflateBatch a, uint32_t *d_counter, cudaStream_t s, int slot, bool may_order)
{
    a.counter = d_counter;
    if (may_order && !a.order && a.n > (uint32_t)ctx->sm_count) {  // heaviest first, so that the longest streams do not start last
        CU(ctx->d_sched[slot].reserve((size_t)a.n * 4));
        dbg::sched_order_kernel<<<1, 1024, 0, s>>>(a, (uint32_t *)ctx->d_sched[slot].p);
        ctx->launches++;
        a.order = (const uint32_t *)ctx->d_sched[slot].p;
    }
    CU(cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), s));
    uint32_t ctas_needed = (a.n + dbg::INFLATE_WARPS_PER_CTA - 1) / dbg::INFLATE_WARPS_PER_CTA;
    uint32_t grid = std::min<uint32_t>(ctas_needed, (uint32_t)ctx->sm_count * ctx->inflate_ctas_per_sm);
    size_t smem = sizeof(dbg::InflateSmem) * dbg::INFLATE_WARPS_PER_CTA;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->profiling) {
        if (ctx->prof_used == ctx->prof_events.size()) {
            CU(cudaEventCreate(&e0));
            CU(cudaEventCreate(&e1));
            ctx->prof_events.push_back({e0, e1});
        }
        e0 = ctx->prof_events[ctx->prof_used].first;
        e1 = ctx->prof_events[ctx->prof_used].second;
        ctx->prof_used++;
        CU(cudaEventRecord(e0, s));
    }
    dbg::inflate_batch_kernel<<<grid, dbg::INFLATE_THREADS, smem, s>>>(a);
    ctx->launches++;
    CU(cudaGetLastError());
    if (e1) CU(cudaEventRecord(e1, s));
    return DBG_OK;
}

// Lane-serial path for streams that are a single fixed-Huffman block (every stb-written PNG): fx_core.h /
// fx_kernels.cuh. Two small device->host reads (how many streams / bytes; exact token and cell counts), so `s`
// is synchronised twice. *skip_out = per-stream flags of the streams handled here, *redo_out = those handed
// back (flag set as well) for a second warp-per-stream pass; *n_redo tells whether there are any.
static int run_fx(dbg_ctx *ctx, const dbg::InflateBatch &a, cudaStream_t s, const uint32_t **skip_out, const uint32_t **redo_out,
                  uint32_t *n_redo)
{
    *skip_out = nullptr;
    *redo_out = nullptr;
    *n_redo = 0;
    const uint32_t n = a.n;
    CU(ctx->h_summary.reserve(sizeof(dbg::FxSummary)));
    CU(ctx->d_split.reserve(256 + (size_t)n * (2 * 8 + 6 * 4) + 256));
    uint8_t *p = (uint8_t *)ctx->d_split.p;
    dbg::FxBatch b{};
    b.in_base = a.in_base; b.in_off = a.in_off; b.in_size = a.in_size;
    b.out_base = a.out_base; b.out_off = a.out_off; b.out_cap = a.out_cap;
    b.out_size = a.out_size; b.status = a.status; b.pre_status = a.pre_status; b.n = n;
    b.summary = (dbg::FxSummary *)p;
    b.cell_base = (uint64_t *)(p + 256);
    b.tok_base = b.cell_base + n;
    b.flag = (uint32_t *)(b.tok_base + n);
    b.redo = b.flag + n;
    b.chunk_base = b.redo + n;
    b.nchunks = b.chunk_base + n;
    b.group_base = b.nchunks + n;
    b.ngroups = b.group_base + n;
    const unsigned sb = (n + 127) / 128;
    CU(cudaMemsetAsync(b.summary, 0, sizeof(dbg::FxSummary), s));
    dbg::fx_classify_kernel<<<sb, 128, 0, s>>>(b);
    ctx->launches++;
    CU(cudaGetLastError());
    dbg::FxSummary *hs = (dbg::FxSummary *)ctx->h_summary.p;
    CU(cudaMemcpyAsync(hs, b.summary, sizeof(dbg::FxSummary), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (hs->n_fx == 0) return DBG_OK;
    // chunk = what one lane decodes: 16 KiB, halved (down to 2 KiB) while the batch has fewer chunks than the
    // GPU has lanes to give; group = what one warp expands and the unit of the marker / resolve scheme:
    // 256 KiB of compressed data, smaller for small batches (more warps), larger for very long streams (the
    // last 32 KiB of every group's output is resolved by a per-stream serial chain).
    const uint64_t lanes_wanted = (uint64_t)ctx->sm_count * 2048;
    uint32_t chunk = 16384;
    while (chunk > dbg::FX_MIN_CHUNK && hs->fx_in / chunk < lanes_wanted) chunk >>= 1;
    if (ctx->split_chunk_forced) chunk = ctx->split_chunk_forced;
    uint32_t group = 262144;
    const uint64_t warps_wanted = (uint64_t)ctx->sm_count * 32;
    while (group > 32768 && hs->fx_in / group < warps_wanted) group >>= 1;
    while (group < (1u << 20) && hs->max_in / group > 384 && hs->fx_in / (2 * group) >= warps_wanted) group <<= 1;
    if (ctx->fx_group_forced) group = ctx->fx_group_forced;
    if (group < chunk) group = chunk;
    b.chunk_bytes = chunk;
    b.group_chunks = group / chunk;
    const uint32_t T = (uint32_t)(hs->fx_in / chunk) + hs->n_fx;               // upper bounds
    const uint32_t NG = (uint32_t)(hs->fx_in / ((uint64_t)chunk * b.group_chunks)) + hs->n_fx;
    b.extra_cap = T / 8 + 1024;
    const size_t per_chunk = (size_t)T * (3 * 32 * 4 + 5 * 4) + (size_t)b.extra_cap * 4 + ((size_t)T + b.extra_cap) * sizeof(dbg::FxRec) +
                             (size_t)NG * (2 * 8 + 4 * 4) + 1024;
    CU(ctx->d_split_chunks.reserve(per_chunk));
    uint8_t *q = (uint8_t *)ctx->d_split_chunks.p;
    b.rec = (dbg::FxRec *)q;
    b.g_out_off = (uint64_t *)(b.rec + T + b.extra_cap);
    b.g_tok_off = b.g_out_off + NG;
    b.hyp = (uint32_t *)(b.g_tok_off + NG);
    b.surv_start = b.hyp + (size_t)T * 32;
    b.extra_slot = b.surv_start + (size_t)T * 32;
    b.chunk_stream = b.extra_slot + (size_t)T * 32;
    b.nsurv = b.chunk_stream + T;
    b.c_surv = b.nsurv + T;
    b.c_out_off = b.c_surv + T;
    b.c_tok_off = b.c_out_off + T;
    b.extra_item = b.c_tok_off + T;
    b.group_stream = b.extra_item + b.extra_cap;
    b.g_out_len = b.group_stream + NG;
    b.g_flag = b.g_out_len + NG;
    b.g_ntok = b.g_flag + NG;
    CU(cudaMemsetAsync(b.chunk_stream, 0, (size_t)T * 2 * 4, s));                    // chunk_stream, nsurv of unused slots
    CU(cudaMemsetAsync(b.c_surv, 0xff, (size_t)T * 4, s));                           // FX_NONE
    CU(cudaMemsetAsync(b.group_stream, 0, (size_t)NG * 4 * 4, s));                   // group_stream, g_out_len, g_flag, g_ntok of unused slots
    const uint32_t warp_grid = std::min<uint32_t>((T + dbg::FX_WARPS_PER_CTA - 1) / dbg::FX_WARPS_PER_CTA, (uint32_t)ctx->sm_count * 12);
    const uint32_t lane_grid = std::min<uint32_t>((T + dbg::FX_LANE_THREADS - 1) / dbg::FX_LANE_THREADS, (uint32_t)ctx->sm_count * 12);
    dbg::fx_assign_kernel<<<sb, 128, 0, s>>>(b);
    dbg::fx_fill_kernel<<<n, 128, 0, s>>>(b);
    dbg::fx_head_kernel<<<warp_grid, dbg::FX_WARPS_PER_CTA * 32, 0, s>>>(b);
    dbg::fx_sizes_kernel<<<lane_grid, dbg::FX_LANE_THREADS, 0, s>>>(b);
    dbg::fx_chain_kernel<<<std::min<uint32_t>((n + dbg::FX_WARPS_PER_CTA - 1) / dbg::FX_WARPS_PER_CTA, (uint32_t)ctx->sm_count * 8),
                           dbg::FX_WARPS_PER_CTA * 32, 0, s>>>(b);
    ctx->launches += 5;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hs, b.summary, sizeof(dbg::FxSummary), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (hs->cells_used) {
        CU(ctx->d_cells.reserve((size_t)hs->cells_used * 2 + 256));
        CU(ctx->d_fx_tok.reserve((size_t)hs->tok_used * 4 + 256));
        b.cells = (uint16_t *)ctx->d_cells.p;
        b.tok = (uint32_t *)ctx->d_fx_tok.p;
        dbg::fx_tokens_kernel<<<lane_grid, dbg::FX_LANE_THREADS, 0, s>>>(b);
        dbg::fx_expand_kernel<<<std::min<uint32_t>((NG + dbg::FX_WARPS_PER_CTA - 1) / dbg::FX_WARPS_PER_CTA, (uint32_t)ctx->sm_count * 12),
                                dbg::FX_WARPS_PER_CTA * 32, 0, s>>>(b);
        // cells -> bytes: the resolve kernels see the groups as their chunks
        dbg::SplitBatch r{};
        r.out_base = a.out_base; r.out_off = a.out_off; r.out_size = a.out_size; r.status = a.status; r.n = n;
        r.split_flag = b.flag; r.redo = b.redo; r.chunk_base = b.group_base; r.nchunks = b.ngroups; r.cell_base = b.cell_base;
        r.chunk_stream = b.group_stream; r.c_out_off = b.g_out_off; r.c_out_len = b.g_out_len; r.c_flag = b.g_flag;
        r.cells = b.cells;
        dbg::split_resolve_tails_kernel<<<n, dbg::RESOLVE_THREADS, 0, s>>>(r);
        dbg::split_resolve_body_kernel<<<std::min<uint32_t>(NG, (uint32_t)ctx->sm_count * 8), 256, 0, s>>>(r, NG);
        ctx->launches += 4;
        CU(cudaGetLastError());
    }
    *skip_out = b.flag;
    *redo_out = b.redo;
    *n_redo = hs->n_redo;
    ctx->fx_streams += hs->n_fx - hs->n_redo;
    ctx->fx_redo += hs->n_redo;
    ctx->fx_extra += hs->n_extra;
    return DBG_OK;
}

static int launch_inflate_plain(dbg_ctx *ctx, dbg::InflateBatch a, uint32_t *d_counter, cudaStream_t s, int slot, bool may_order = true);

// When streams do take this path, the warp-per-stream kernel for all the others is launched from here, on an
// auxiliary stream, as soon as the classification is known: it is bound by the latency of its longest streams
// and leaves most SM slots free, which the (throughput-bound) block-split kernels then fill. *regular_done
// tells the caller that this has happened.
static int run_bsplit(dbg_ctx *ctx, int slot, dbg::InflateBatch a, uint32_t *d_counter, cudaStream_t s, const uint32_t *taken,
                      bool *regular_done)
{
    *regular_done = false;
    const uint32_t n = a.n;
    Buf &bs = ctx->d_bs_stream[slot], &br = ctx->d_bs_region[slot], &bc = ctx->d_bs_cells[slot], &hsb = ctx->h_bs_summary[slot];
    CU(hsb.reserve(sizeof(dbg::BsSummary)));
    CU(bs.reserve(256 + (size_t)n * (8 + 8 + 4 + 4 + 4 + 4) + 256));
    uint8_t *p = (uint8_t *)bs.p;
    dbg::BsBatch b{};
    b.in_base = a.in_base; b.in_off = a.in_off; b.in_size = a.in_size;
    b.out_base = a.out_base; b.out_off = a.out_off; b.out_cap = a.out_cap;
    b.out_size = a.out_size; b.status = a.status; b.pre_status = a.pre_status; b.taken = taken; b.n = n;
    b.resident_warps = (uint32_t)ctx->sm_count * ctx->inflate_ctas_per_sm * dbg::INFLATE_WARPS_PER_CTA;
    b.min_bytes = ctx->bsplit_min_bytes;
    b.factor_q = ctx->bsplit_factor_q;
    b.summary = (dbg::BsSummary *)p;
    b.cell_base = (uint64_t *)(p + 256);
    b.tok_stream_base = b.cell_base + n;
    b.flag = (uint32_t *)(b.tok_stream_base + n);
    b.chunk_base = b.flag + n;
    b.nchunks = b.chunk_base + n;
    b.redo = b.nchunks + n;
    const unsigned sb = (n + 127) / 128;
    CU(cudaMemsetAsync(b.summary, 0, sizeof(dbg::BsSummary), s));
    dbg::bs_sum_kernel<<<sb, 128, 0, s>>>(b);
    dbg::bs_classify_kernel<<<sb, 128, 0, s>>>(b);
    ctx->launches += 2;
    CU(cudaGetLastError());
    dbg::BsSummary *hs = (dbg::BsSummary *)hsb.p;
    CU(cudaMemcpyAsync(hs, b.summary, sizeof(dbg::BsSummary), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (hs->n_split == 0) return DBG_OK;
    // fork: everything that is not split, on the auxiliary stream
    a.skip = taken;
    a.skip2 = b.flag;
    CU(cudaEventRecord(ctx->aux_fork, s));
    CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
    {
        int rc = launch_inflate_plain(ctx, a, d_counter, ctx->aux_stream, slot);
        if (rc) return rc;
    }
    CU(cudaEventRecord(ctx->aux_join, ctx->aux_stream));
    *regular_done = true;
    // region size: 64 KiB when that already gives every resident warp a few regions, else smaller (more, shorter
    // chunks: the latency of a lone long stream is the decode time of its longest chunk)
    uint32_t region = ctx->bsplit_region;
    while (region > ctx->bsplit_region_min && hs->split_in / region < 4ull * b.resident_warps) region >>= 1;
    b.region_bytes = region;
    const uint32_t T = (uint32_t)(hs->split_in / region) + hs->n_split;  // upper bound of the region count
    CU(br.reserve((size_t)T * (4 + 8 + 8 + 8 + 4 + 4 + 4) + 256));
    CU(cudaMemsetAsync(br.p, 0, (size_t)T * (4 + 8 + 8 + 8 + 4 + 4 + 4), s));
    b.cand = (uint64_t *)br.p;
    b.exit_bits = b.cand + T;
    b.c_out_off = b.exit_bits + T;
    b.chunk_stream = (uint32_t *)(b.c_out_off + T);
    b.c_out_len = b.chunk_stream + T;
    b.c_flag = b.c_out_len + T;
    b.c_ntok = b.c_flag + T;
    // token areas: the count pass records every symbol so that the second pass need not decode Huffman codes again
    uint32_t tpb = ctx->bsplit_tok_per_byte;
    while (tpb > 1 && (uint64_t)tpb * 4 * hs->split_in > ctx->bsplit_tok_max_bytes) tpb >>= 1;
    if (tpb && (uint64_t)tpb * 4 * hs->split_in <= ctx->bsplit_tok_max_bytes) {
        if (ctx->d_bs_tok[slot].reserve((size_t)tpb * 4 * hs->split_in + 256) == cudaSuccess) {
            b.tok = (uint32_t *)ctx->d_bs_tok[slot].p;
            b.tok_per_byte = tpb;
        } else {
            (void)cudaGetLastError();  // no room for tokens: the second pass decodes the Huffman codes again
        }
    }
    const size_t smem = sizeof(dbg::InflateSmem) * dbg::BS_WARPS_PER_CTA;
    const uint32_t grid = std::min<uint32_t>((T + dbg::BS_WARPS_PER_CTA - 1) / dbg::BS_WARPS_PER_CTA,
                                             (uint32_t)ctx->sm_count * dbg::INFLATE_CTAS_PER_SM);
    dbg::bs_assign_kernel<<<sb, 128, 0, s>>>(b);
    ctx->launches++;
    dbg::bs_fill_kernel<<<n, 128, 0, s>>>(b);
    dbg::bs_search_kernel<<<grid, dbg::BS_WARPS_PER_CTA * 32, 0, s>>>(b);
    dbg::bs_prune_kernel<<<sb, 128, 0, s>>>(b);
    ctx->launches += 2;
    dbg::bs_count_kernel<<<grid, dbg::BS_WARPS_PER_CTA * 32, smem, s>>>(b);
    dbg::bs_chain_kernel<<<sb, 128, 0, s>>>(b);
    ctx->launches += 3;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hs, b.summary, sizeof(dbg::BsSummary), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    ctx->bs_streams += hs->n_split - hs->n_pruned - hs->n_fallback;
    ctx->bs_fallbacks += hs->n_fallback;
    if (hs->cells_used) {
        CU(bc.reserve((size_t)hs->cells_used * 2 + 256));
        b.cells = (uint16_t *)bc.p;
        dbg::bs_decode_kernel<<<grid, dbg::BS_WARPS_PER_CTA * 32, smem, s>>>(b);
        // cells -> bytes with the split-stream path's resolve kernels
        dbg::SplitBatch r{};
        r.out_base = a.out_base; r.out_off = a.out_off; r.out_size = a.out_size; r.status = a.status; r.n = n;
        r.split_flag = b.flag; r.redo = b.redo; r.chunk_base = b.chunk_base; r.nchunks = b.nchunks; r.cell_base = b.cell_base;
        r.chunk_stream = b.chunk_stream; r.c_out_off = b.c_out_off; r.c_out_len = b.c_out_len; r.c_flag = b.c_flag;
        r.cells = b.cells;
        dbg::split_resolve_tails_kernel<<<n, dbg::RESOLVE_THREADS, 0, s>>>(r);
        dbg::split_resolve_body_kernel<<<std::min<uint32_t>(T, (uint32_t)ctx->sm_count * 8), 256, 0, s>>>(r, T);
        ctx->launches += 3;
        CU(cudaGetLastError());
    }
    if (hs->n_fallback || hs->n_pruned) {
        // second warp-per-stream pass for the streams whose hinted boundaries did not chain or were too sparse
        dbg::InflateBatch again = a;
        again.skip = nullptr;
        again.skip2 = nullptr;
        again.only = b.redo;
        again.order = nullptr;  // and no device-made order either: the first pass may still be reading that buffer
        int rc = launch_inflate_plain(ctx, again, d_counter + 1, s, slot, false);
        if (rc) return rc;
    }
    CU(cudaStreamWaitEvent(s, ctx->aux_join, 0));  // join
    return DBG_OK;
}

static int launch_inflate(dbg_ctx *ctx, dbg::InflateBatch a, uint32_t *d_counter, cudaStream_t s, int slot = 0)
{
    a.skip = nullptr;
    a.skip2 = nullptr;
    const uint32_t *fx_redo = nullptr;
    uint32_t n_redo = 0;
    if (ctx->fx && slot == 0 && ctx->bsplit_allowed) {  // (the scratch of this path is per context: single-wave calls only)
        const uint32_t *skip = nullptr;
        int rc = run_fx(ctx, a, s, &skip, &fx_redo, &n_redo);
        if (rc) return rc;
        a.skip = skip;
    }
    if (n_redo) {
        // streams the lane-serial path handed back: a warp-per-stream pass of their own (a.skip keeps them out of
        // everything below)
        dbg::InflateBatch again = a;
        again.skip = nullptr;
        again.only = fx_redo;
        again.order = nullptr;
        int rc = launch_inflate_plain(ctx, again, d_counter + 2, s, slot, false);
        if (rc) return rc;
    }
    if (ctx->bsplit && ctx->bsplit_allowed) {
        bool done = false;
        int rc = run_bsplit(ctx, slot, a, d_counter, s, a.skip, &done);
        if (rc || done) return rc;
    }
    return launch_inflate_plain(ctx, a, d_counter, s, slot);
}

static int inflate_device_slot(dbg_ctx *ctx, int slot, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                               const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                               const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_status,
                               const uint32_t *d_order, cudaStream_t s, uint64_t *gz_off, uint64_t *gz_size,
                               uint32_t *gz_pre)
{
    CU(ctx->d_counter.reserve(8 * dbg_ctx::MAX_WAVES * sizeof(uint32_t)));
    uint32_t *counter = (uint32_t *)ctx->d_counter.p + 8 * slot;
    if (gz_off) {
        dbg::gz_scan_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, d_in_off, d_in_size, (uint32_t)n, gz_off, gz_size, gz_pre);
        ctx->launches++;
        CU(cudaGetLastError());
        dbg::InflateBatch a{d_in, gz_off, gz_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, gz_pre, nullptr, nullptr, nullptr, d_order, nullptr, (uint32_t)n};
        int rc = launch_inflate(ctx, a, counter, s, slot);
        if (rc || !ctx->verify) return rc;
        uint32_t ctas = (uint32_t)std::min<uint64_t>((n + dbg::SCAN_WARPS - 1) / dbg::SCAN_WARPS, (uint64_t)ctx->sm_count * 8);
        dbg::gz_verify_kernel<<<ctas, dbg::SCAN_WARPS * 32, 0, s>>>(d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_size, d_status,
                                                                    (uint32_t)n);
        ctx->launches++;
        CU(cudaGetLastError());
        return DBG_OK;
    }
    dbg::InflateBatch a{d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, nullptr, nullptr, nullptr, nullptr, d_order, nullptr, (uint32_t)n};
    return launch_inflate(ctx, a, counter, s, slot);
}

extern "C" int dbg_inflate_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                                        const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                                        const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_status,
                                        const uint32_t *d_order, void *stream)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!d_in || !d_in_off || !d_in_size || !d_out || !d_out_off || !d_out_cap || !d_out_size || !d_status ||
        n > 0x7fffffffull) {
        set_err(ctx, "dbg_inflate_batch_device: bad arguments");
        return DBG_ERR_ARG;
    }
    ctx->bsplit_allowed = true;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    return inflate_device_slot(ctx, 0, n, d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, d_order,
                               s, nullptr, nullptr, nullptr);
}

extern "C" int dbg_decode_gz_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                                          const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                                          const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_status,
                                          const uint32_t *d_order, void *stream)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!d_in || !d_in_off || !d_in_size || !d_out || !d_out_off || !d_out_cap || !d_out_size || !d_status ||
        n > 0x7fffffffull) {
        set_err(ctx, "dbg_decode_gz_batch_device: bad arguments");
        return DBG_ERR_ARG;
    }
    ctx->bsplit_allowed = true;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    CU(ctx->d_meta.reserve(n * 20 + 64));
    uint64_t *p_off = (uint64_t *)ctx->d_meta.p;
    uint64_t *p_size = p_off + n;
    uint32_t *pre = (uint32_t *)(p_size + n);
    return inflate_device_slot(ctx, 0, n, d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, d_order,
                               s, p_off, p_size, pre);
}

extern "C" uint64_t dbg_png_scratch_bytes(uint64_t n, uint64_t total_in_bytes, uint64_t total_rgba_bytes)
{
    return dbg::png_scratch_bytes(n, total_in_bytes, total_rgba_bytes);
}

extern "C" int dbg_decode_png_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                                           const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                                           const uint64_t *d_out_cap, uint32_t *d_status, uint64_t total_in_bytes,
                                           uint64_t total_rgba_bytes, void *stream)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!d_in || !d_in_off || !d_in_size || !d_out || !d_out_off || !d_out_cap || !d_status || n > 0x7fffffffull) {
        set_err(ctx, "dbg_decode_png_batch_device: bad arguments");
        return DBG_ERR_ARG;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    CU(ctx->d_counter.reserve(8 * dbg_ctx::MAX_WAVES * sizeof(uint32_t)));
    CU(ctx->d_png_scratch.reserve(dbg::png_scratch_bytes(n, total_in_bytes, total_rgba_bytes)));
    dbg::PngLayout lay = dbg::png_layout((uint8_t *)ctx->d_png_scratch.p, n, total_in_bytes, total_rgba_bytes);

    // 1. container walk + CRC-32 + IDAT gather (one warp per image)
    dbg::PngBatch pb{d_in, d_in_off, d_in_size, d_out_cap, (uint32_t)n, lay};
    int rc = dbg::png_launch_scan(pb, ctx->sm_count, s);
    ctx->launches += 4;
    if (rc) CU((cudaError_t)rc);
    // 2. inflate the compacted zlib payloads into the filtered-scanline buffers
    // z_off holds absolute addresses: a single-IDAT image is inflated straight from the file
    dbg::InflateBatch a{nullptr, lay.z_off, lay.z_size, lay.scan, lay.s_off, lay.s_cap, lay.s_size, lay.inf_status,
                        lay.pre_status, nullptr, nullptr, nullptr, nullptr, nullptr, (uint32_t)n};
    rc = launch_inflate(ctx, a, (uint32_t *)ctx->d_counter.p, s);
    if (rc) return rc;
    if (ctx->verify) {
        uint32_t ctas = (uint32_t)std::min<uint64_t>((n + dbg::SCAN_WARPS - 1) / dbg::SCAN_WARPS, (uint64_t)ctx->sm_count * 8);
        dbg::png_adler_kernel<<<ctas, dbg::SCAN_WARPS * 32, 0, s>>>(pb);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    // 3. un-filter (+ palette / RGB expansion) straight into the caller's RGBA
    rc = dbg::png_launch_unfilter(pb, d_out, d_out_off, d_status, ctx->sm_count, s);
    ctx->launches += 2;
    if (rc) CU((cudaError_t)rc);
    return DBG_OK;
}

// -------------------------------------------------------------- host batches --
static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// Scheduling weight of an item for the largest-first work queue: its compressed size, except that a
// stream whose first block is stored (a plain copy, ~16x cheaper per byte than Huffman decode) counts less.
static inline uint64_t sched_weight(int kind, const uint8_t *p, uint64_t size)
{
    uint64_t at = 0;
    if (kind == 1) {  // gzip: skip the header the way gz_scan_kernel does
        if (size < 18) return size;
        at = 10;
        if ((p[3] >> 3) & 1) {
            while (at < size && p[at] != 0) at++;
            at++;
        }
    }
    if (kind == 2 || at >= size) return size;
    return ((p[at] >> 1) & 3) == 0 ? size / 64 : size;
}

extern "C" int dbg_decode_batch_packed(dbg_ctx *ctx, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                                       const uint64_t *in_size, uint8_t *h_out, const uint64_t *out_off,
                                       const uint64_t *out_cap, uint64_t *out_size, uint32_t *status)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!h_in || !in_off || !in_size || !h_out || !out_off || !out_cap || !out_size || !status || kind < 0 || kind > 2 ||
        n > 0x7fffffffull) {
        set_err(ctx, "dbg_decode_batch_packed: bad arguments");
        return DBG_ERR_ARG;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    uint64_t in_span = 0, out_span = 0, tot_in = 0, tot_out = 0;
    for (uint64_t i = 0; i < n; i++) {
        in_span = std::max(in_span, in_off[i] + in_size[i]);
        out_span = std::max(out_span, out_off[i] + out_cap[i]);
        tot_in += in_size[i];
        tot_out += out_cap[i];
    }
    // descriptor block: in_off, in_size, out_off, out_cap, out_size (u64 x n each), status + order (u32 x n each)
    size_t desc_bytes = n * (5 * 8 + 2 * 4);
    CU(ctx->h_desc.reserve(desc_bytes));
    CU(ctx->d_desc.reserve(desc_bytes));
    CU(ctx->d_in.reserve(in_span + 64));
    CU(ctx->d_out.reserve(out_span + 64));
    uint64_t *hd = (uint64_t *)ctx->h_desc.p;
    memcpy(hd, in_off, n * 8);
    memcpy(hd + n, in_size, n * 8);
    memcpy(hd + 2 * n, out_off, n * 8);
    memcpy(hd + 3 * n, out_cap, n * 8);
    uint32_t *h_status = (uint32_t *)(hd + 5 * n);
    uint32_t *h_order = h_status + n;
    uint64_t *dd = (uint64_t *)ctx->d_desc.p;
    uint32_t *d_status = (uint32_t *)(dd + 5 * n);
    uint32_t *d_order = d_status + n;
    // ---- waves: when the items sit in index order in both arenas, the batch is cut into up to
    // MAX_WAVES contiguous waves, each on its own stream (H2D -> kernels -> D2H), so the copy
    // engines and the SMs overlap. Otherwise (or for PNG, whose scratch is per batch) one wave.
    bool mono = kind != 2;
    for (uint64_t i = 1; i < n && mono; i++)
        mono = in_off[i] >= in_off[i - 1] + in_size[i - 1] && out_off[i] >= out_off[i - 1] + out_cap[i - 1];
    int nw = mono ? (int)std::min<uint64_t>(ctx->waves, std::max<uint64_t>(1, n / 256)) : 1;
    if (n < ctx->split_max_streams) nw = 1;  // small batches may take the split-stream path, which owns per-context scratch
    if (nw > 1 && ctx->bsplit) {
        // a batch with streams long enough for the block-split path (same rule as bs_classify_kernel) runs
        // as one wave: that path synchronises the host twice, which would serialise concurrent waves
        const uint64_t resident = (uint64_t)ctx->sm_count * ctx->inflate_ctas_per_sm * dbg::INFLATE_WARPS_PER_CTA;
        const uint64_t thr = std::max<uint64_t>(ctx->bsplit_min_bytes, tot_in / resident * ctx->bsplit_factor_q / 4);
        for (uint64_t i = 0; i < n && nw > 1; i++)
            if (in_size[i] >= thr) nw = 1;
    }
    ctx->bsplit_allowed = nw == 1;
    CU(cudaMemcpyAsync(dd, hd, 4 * n * 8, cudaMemcpyHostToDevice, s));
    if (kind == 1) CU(ctx->d_meta.reserve(n * 20 + 64));
    uint64_t *gz_off = kind == 1 ? (uint64_t *)ctx->d_meta.p : nullptr;
    uint64_t *gz_size = gz_off ? gz_off + n : nullptr;
    uint32_t *gz_pre = gz_off ? (uint32_t *)(gz_size + n) : nullptr;
    if (nw == 1) {
        std::vector<uint64_t> wt(n);
        for (uint64_t i = 0; i < n; i++) wt[i] = sched_weight(kind, h_in + in_off[i], in_size[i]) + out_cap[i] / 32;
        std::iota(h_order, h_order + n, 0u);
        std::stable_sort(h_order, h_order + n, [&](uint32_t a, uint32_t b) { return wt[a] > wt[b]; });
        CU(cudaMemcpyAsync(d_order, h_order, n * 4, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(ctx->d_in.p, h_in, in_span, cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync((uint8_t *)ctx->d_in.p + in_span, 0, 64, s));
        int rc;
        if (kind == 2)
            rc = dbg_decode_png_batch_device(ctx, n, (const uint8_t *)ctx->d_in.p, dd, dd + n, (uint8_t *)ctx->d_out.p,
                                             dd + 2 * n, dd + 3 * n, d_status, tot_in, tot_out, s);
        else
            rc = inflate_device_slot(ctx, 0, n, (const uint8_t *)ctx->d_in.p, dd, dd + n, (uint8_t *)ctx->d_out.p, dd + 2 * n,
                                     dd + 3 * n, dd + 4 * n, d_status, d_order, s, gz_off, gz_size, gz_pre);
        if (rc) return rc;
        CU(cudaMemcpyAsync(h_out, ctx->d_out.p, out_span, cudaMemcpyDeviceToHost, s));
    } else {
        // per-wave largest-first order (indices relative to the wave)
        std::vector<uint64_t> cut(nw + 1);
        for (int k = 0; k <= nw; k++) cut[k] = n * (uint64_t)k / nw;
        for (int k = 0; k < nw; k++) {
            uint32_t *o = h_order + cut[k];
            uint64_t m = cut[k + 1] - cut[k], b = cut[k];
            std::vector<uint64_t> wt(m);
            for (uint64_t i = 0; i < m; i++) wt[i] = sched_weight(kind, h_in + in_off[b + i], in_size[b + i]) + out_cap[b + i] / 32;
            std::iota(o, o + m, 0u);
            std::stable_sort(o, o + m, [&](uint32_t x, uint32_t y) { return wt[x] > wt[y]; });
        }
        CU(cudaMemcpyAsync(d_order, h_order, n * 4, cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync((uint8_t *)ctx->d_in.p + in_span, 0, 64, s));
        CU(cudaEventRecord(ctx->wave_ready, s));
        const bool trace = getenv("DBG_WAVE_TRACE") != nullptr;  // debugging aid: per-wave timeline on stderr
        std::vector<cudaEvent_t> tev;
        if (trace) {
            tev.resize(3 * nw + 1);
            for (auto &e : tev) cudaEventCreate(&e);
            cudaEventRecord(tev[3 * nw], s);
        }
        // Issued breadth-first (all uploads, then all kernels, then all downloads): the streams share a limited
        // number of hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default), and with wave-by-wave issue the
        // upload of wave k+8 sits behind the download of wave k, i.e. behind wave k's kernels (measured with
        // DBG_WAVE_TRACE: the second half of the waves did not start before the first half had finished).
        for (int k = 0; k < nw; k++) {
            cudaStream_t ws = ctx->wave_stream[k];
            uint64_t b = cut[k], e = cut[k + 1];
            uint64_t i0 = in_off[b], i1 = in_off[e - 1] + in_size[e - 1];
            CU(cudaStreamWaitEvent(ws, ctx->wave_ready, 0));
            CU(cudaMemcpyAsync((uint8_t *)ctx->d_in.p + i0, h_in + i0, i1 - i0, cudaMemcpyHostToDevice, ws));
            if (trace) cudaEventRecord(tev[3 * k], ws);
        }
        for (int k = 0; k < nw; k++) {
            cudaStream_t ws = ctx->wave_stream[k];
            uint64_t b = cut[k], e = cut[k + 1], m = e - b;
            int rc = inflate_device_slot(ctx, k, m, (const uint8_t *)ctx->d_in.p, dd + b, dd + n + b, (uint8_t *)ctx->d_out.p,
                                         dd + 2 * n + b, dd + 3 * n + b, dd + 4 * n + b, d_status + b, d_order + b, ws,
                                         gz_off ? gz_off + b : nullptr, gz_off ? gz_size + b : nullptr,
                                         gz_off ? gz_pre + b : nullptr);
            if (rc) return rc;
            if (trace) cudaEventRecord(tev[3 * k + 1], ws);
        }
        for (int k = 0; k < nw; k++) {
            cudaStream_t ws = ctx->wave_stream[k];
            uint64_t b = cut[k], e = cut[k + 1];
            uint64_t o0 = out_off[b], o1 = out_off[e - 1] + out_cap[e - 1];
            CU(cudaMemcpyAsync(h_out + o0, (uint8_t *)ctx->d_out.p + o0, o1 - o0, cudaMemcpyDeviceToHost, ws));
            if (trace) cudaEventRecord(tev[3 * k + 2], ws);
        }
        for (int k = 0; k < nw; k++) CU(cudaStreamSynchronize(ctx->wave_stream[k]));
        if (trace) {
            for (int k = 0; k < nw; k++) {
                float a = 0, b2 = 0, c = 0;
                cudaEventElapsedTime(&a, tev[3 * nw], tev[3 * k]);
                cudaEventElapsedTime(&b2, tev[3 * nw], tev[3 * k + 1]);
                cudaEventElapsedTime(&c, tev[3 * nw], tev[3 * k + 2]);
                fprintf(stderr, "wave %2d: h2d done %7.2f ms, kernels done %7.2f ms, d2h done %7.2f ms\n", k, a, b2, c);
            }
            for (auto &e : tev) cudaEventDestroy(e);
        }
    }
    ctx->bsplit_allowed = true;
    CU(cudaMemcpyAsync(hd + 4 * n, dd + 4 * n, n * 8 + n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(status, h_status, n * 4);
    if (kind == 2) {
        for (uint64_t i = 0; i < n; i++) out_size[i] = status[i] == 0 ? out_cap[i] : 0;
    } else {
        memcpy(out_size, hd + 4 * n, n * 8);
    }
    return DBG_OK;
}

// pointer-array front ends: pack into the pinned staging arenas, run the packed
// path, scatter the results.
static int run_pointer_batch(dbg_ctx *ctx, int kind, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                             uint8_t *const *out, const uint64_t *out_cap, uint64_t *out_size, uint32_t *status)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!in || !in_size || !out || !out_cap || !status) {
        set_err(ctx, "batch call: NULL argument");
        return DBG_ERR_ARG;
    }
    CU(cudaSetDevice(ctx->device));
    std::vector<uint64_t> in_off(n), out_off(n), sizes(n);
    uint64_t ti = 0, to = 0;
    for (uint64_t i = 0; i < n; i++) {
        in_off[i] = ti;
        ti += align_up(in_size[i] + 16, 16);
        out_off[i] = to;
        to += align_up(out_cap[i], 16);
    }
    CU(ctx->h_in.reserve(ti + 64));
    CU(ctx->h_out.reserve(to + 64));
    uint8_t *hi = (uint8_t *)ctx->h_in.p;
    for (uint64_t i = 0; i < n; i++) {
        if (in_size[i] && !in[i]) {
            set_err(ctx, "batch call: item %llu has a NULL input", (unsigned long long)i);
            return DBG_ERR_ARG;
        }
        memcpy(hi + in_off[i], in[i], in_size[i]);
        memset(hi + in_off[i] + in_size[i], 0, align_up(in_size[i] + 16, 16) - in_size[i]);
    }
    int rc = dbg_decode_batch_packed(ctx, kind, n, hi, in_off.data(), in_size, (uint8_t *)ctx->h_out.p, out_off.data(),
                                     out_cap, sizes.data(), status);
    if (rc) return rc;
    for (uint64_t i = 0; i < n; i++) {
        if (status[i] == 0 && out[i]) memcpy(out[i], (uint8_t *)ctx->h_out.p + out_off[i], sizes[i]);
        if (out_size) out_size[i] = sizes[i];
    }
    return DBG_OK;
}

extern "C" int dbg_inflate_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                                 uint8_t *const *out, const uint64_t *out_cap, uint64_t *out_size, uint32_t *good)
{
    int rc = run_pointer_batch(ctx, 0, n, in, in_size, out, out_cap, out_size, good);
    if (rc == DBG_OK)
        for (uint64_t i = 0; i < n; i++) good[i] = good[i] == 0 ? 1u : 0u;
    return rc;
}

extern "C" int dbg_decode_gz_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                                   uint8_t *const *out, const uint64_t *out_cap, uint64_t *out_size, uint32_t *good)
{
    int rc = run_pointer_batch(ctx, 1, n, in, in_size, out, out_cap, out_size, good);
    if (rc == DBG_OK)
        for (uint64_t i = 0; i < n; i++) good[i] = good[i] == 0 ? 1u : 0u;
    return rc;
}

extern "C" int dbg_decode_png_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                                    uint8_t *const *out_rgba, const uint64_t *rgba_size, uint8_t *good)
{
    if (!good) return DBG_ERR_ARG;
    std::vector<uint32_t> st(n);
    int rc = run_pointer_batch(ctx, 2, n, in, in_size, out_rgba, rgba_size, nullptr, st.data());
    if (rc == DBG_OK)
        for (uint64_t i = 0; i < n; i++) good[i] = st[i] == 0 ? 1 : 0;
    return rc;
}

// ----------------------------------------------------------------------- BMP --
// decode_bmp.c:105-372: header checks + BGRA <-> RGBA swizzle (bmp_kernels.cuh).
static int bmp_launch(dbg_ctx *ctx, bool encode, dbg::BmpBatch b, cudaStream_t s)
{
    const uint64_t n = b.n;
    CU(ctx->d_meta.reserve(n * sizeof(dbg::BmpItem) + (n + 1) * sizeof(uint32_t) + 64));
    b.items = (dbg::BmpItem *)ctx->d_meta.p;
    b.tile_base = (uint32_t *)(b.items + n);
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (encode) dbg::bmp_encode_plan_kernel<<<blocks, 128, 0, s>>>(b);
    else dbg::bmp_decode_plan_kernel<<<blocks, 128, 0, s>>>(b);
    dbg::bmp_scan_kernel<<<1, 1024, 0, s>>>(b.tile_base, b.n);
    dbg::bmp_swizzle_kernel<<<(unsigned)ctx->sm_count * 8, dbg::BMP_THREADS, 0, s>>>(b);
    ctx->launches += 3;
    CU(cudaGetLastError());
    return DBG_OK;
}

extern "C" int dbg_decode_bmp_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                                           const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                                           const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_width,
                                           uint32_t *d_height, uint32_t *d_status, void *stream)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!d_in || !d_in_off || !d_in_size || !d_out || !d_out_off || !d_out_cap || !d_out_size || !d_status ||
        n > 0x7fffffffull) {
        set_err(ctx, "dbg_decode_bmp_batch_device: bad arguments");
        return DBG_ERR_ARG;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    dbg::BmpBatch b{d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, d_width, d_height,
                    (uint32_t)n, nullptr, nullptr};
    return bmp_launch(ctx, false, b, s);
}

extern "C" int dbg_encode_bmp_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_rgba, const uint64_t *d_rgba_off,
                                           const uint64_t *d_rgba_size, const uint32_t *d_width, const uint32_t *d_height,
                                           uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                                           uint64_t *d_out_size, uint32_t *d_status, void *stream)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!d_rgba || !d_rgba_off || !d_rgba_size || !d_width || !d_height || !d_out || !d_out_off || !d_out_cap || !d_out_size ||
        !d_status || n > 0x7fffffffull) {
        set_err(ctx, "dbg_encode_bmp_batch_device: bad arguments");
        return DBG_ERR_ARG;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    dbg::BmpBatch b{d_rgba, d_rgba_off, d_rgba_size, d_out, d_out_off, d_out_cap, d_out_size, d_status,
                    const_cast<uint32_t *>(d_width), const_cast<uint32_t *>(d_height), (uint32_t)n, nullptr, nullptr};
    return bmp_launch(ctx, true, b, s);
}

// Host-pointer front end shared by decode and encode: stage, run, scatter.
static int bmp_host_batch(dbg_ctx *ctx, bool encode, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                          uint32_t *width, uint32_t *height, uint8_t *const *out, const uint64_t *out_cap,
                          uint64_t *out_size, uint32_t *status)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!in || !in_size || !out || !out_cap || !status || (encode && (!width || !height)) || n > 0x7fffffffull) {
        set_err(ctx, "BMP batch call: bad arguments");
        return DBG_ERR_ARG;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::vector<uint64_t> in_off(n), out_off(n);
    uint64_t ti = 0, to = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (in_size[i] && !in[i]) {
            set_err(ctx, "BMP batch call: item %llu has a NULL input", (unsigned long long)i);
            return DBG_ERR_ARG;
        }
        in_off[i] = ti;
        ti += align_up(in_size[i] + 16, 16);
        out_off[i] = to;
        to += align_up(out_cap[i], 16);
    }
    // descriptors: in_off, in_size, out_off, out_cap, out_size (u64) | status, width, height (u32)
    const size_t desc_bytes = n * (5 * 8 + 3 * 4);
    CU(ctx->h_desc.reserve(desc_bytes));
    CU(ctx->d_desc.reserve(desc_bytes));
    CU(ctx->h_in.reserve(ti + 64));
    CU(ctx->d_in.reserve(ti + 64));
    CU(ctx->h_out.reserve(to + 64));
    CU(ctx->d_out.reserve(to + 64));
    uint64_t *hd = (uint64_t *)ctx->h_desc.p, *dd = (uint64_t *)ctx->d_desc.p;
    uint32_t *h32 = (uint32_t *)(hd + 5 * n), *d32 = (uint32_t *)(dd + 5 * n);
    memcpy(hd, in_off.data(), n * 8);
    memcpy(hd + n, in_size, n * 8);
    memcpy(hd + 2 * n, out_off.data(), n * 8);
    memcpy(hd + 3 * n, out_cap, n * 8);
    if (encode) {
        memcpy(h32 + n, width, n * 4);
        memcpy(h32 + 2 * n, height, n * 4);
    }
    uint8_t *hi = (uint8_t *)ctx->h_in.p;
    for (uint64_t i = 0; i < n; i++) memcpy(hi + in_off[i], in[i], in_size[i]);
    CU(cudaMemcpyAsync(dd, hd, desc_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->d_in.p, hi, ti, cudaMemcpyHostToDevice, s));
    int rc;
    if (encode)
        rc = dbg_encode_bmp_batch_device(ctx, n, (const uint8_t *)ctx->d_in.p, dd, dd + n, d32 + n, d32 + 2 * n,
                                         (uint8_t *)ctx->d_out.p, dd + 2 * n, dd + 3 * n, dd + 4 * n, d32, s);
    else
        rc = dbg_decode_bmp_batch_device(ctx, n, (const uint8_t *)ctx->d_in.p, dd, dd + n, (uint8_t *)ctx->d_out.p, dd + 2 * n,
                                         dd + 3 * n, dd + 4 * n, d32 + n, d32 + 2 * n, d32, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(hd + 4 * n, dd + 4 * n, n * 8 + 3 * n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, to, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (uint64_t i = 0; i < n; i++) {
        status[i] = h32[i];
        const uint64_t sz = hd[4 * n + i];
        // encode reports one byte more than it writes (decode_bmp.c:311)
        const uint64_t written = encode && sz ? sz - 1 : sz;
        if (status[i] == 0 && out[i]) memcpy(out[i], (uint8_t *)ctx->h_out.p + out_off[i], written);
        if (out_size) out_size[i] = sz;
        if (!encode) {
            if (width) width[i] = h32[n + i];
            if (height) height[i] = h32[2 * n + i];
        }
    }
    return DBG_OK;
}

extern "C" int dbg_decode_bmp_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                                    uint8_t *const *out_rgba, const uint64_t *rgba_cap, uint32_t *width, uint32_t *height,
                                    uint8_t *good)
{
    if (!good) return DBG_ERR_ARG;
    std::vector<uint32_t> st(n);
    int rc = bmp_host_batch(ctx, false, n, in, in_size, width, height, out_rgba, rgba_cap, nullptr, st.data());
    if (rc == DBG_OK)
        for (uint64_t i = 0; i < n; i++) good[i] = st[i] == 0 ? 1 : 0;
    return rc;
}

extern "C" int dbg_encode_bmp_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *rgba, const uint64_t *rgba_size,
                                    const uint32_t *width, const uint32_t *height, uint8_t *const *out, const uint64_t *out_cap,
                                    uint64_t *out_size, uint32_t *status)
{
    return bmp_host_batch(ctx, true, n, rgba, rgba_size, const_cast<uint32_t *>(width), const_cast<uint32_t *>(height), out,
                          out_cap, out_size, status);
}
