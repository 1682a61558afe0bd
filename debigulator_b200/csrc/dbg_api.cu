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
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <numeric>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/debigulator_b200.h"
#include "bmp_kernels.cuh"
#include "bsplit_kernels.cuh"
#include "fx_kernels.cuh"
#include "kernels.cuh"
#include "png_kernels.cuh"
#include "split_kernels.cuh"
#include "sprite_kernels.cuh"

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

// Scratch of one batch in flight. The device-resident entry points use slot 0; the packed host API cuts a batch
// into waves and gives every wave its own slot and stream, so that the waves' uploads, kernels and downloads overlap.
struct Slot {
    cudaStream_t stream = nullptr;   // the wave's stream
    cudaEvent_t done = nullptr;      // recorded behind the last work that used this slot's scratch
    cudaEvent_t uploaded = nullptr;  // packed host API: the wave's input has arrived
    cudaStream_t aux_stream = nullptr;  // the warp-per-stream kernel runs here beside the block-split kernels
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    Buf counter;                     // work-queue heads
    Buf meta;                        // derived descriptors (gzip payloads, BMP items)
    Buf png_scratch;                 // PNG: per-image meta, task queues, compacted IDAT, filtered scanlines
    Buf sched;                       // work-queue order computed on the device when the caller brings none
    Buf round_tok;                   // token scratch of the warp-per-stream kernel's lane-parallel rounds (64 KiB per resident warp)
    Buf fx_stream, fx_chunks, fx_tok, cells, h_fx;            // lane-serial fixed-block path (fx_kernels.cuh)
    Buf bs_stream, bs_region, bs_cells, bs_tok, bs_pieces, h_bs;  // block-split path (bsplit_kernels.cuh)
    Slot() { h_fx.pinned_host = h_bs.pinned_host = true; }
    void release()
    {
        Buf *all[] = {&counter, &meta, &png_scratch, &sched, &round_tok, &fx_stream, &fx_chunks, &fx_tok, &cells, &h_fx,
                      &bs_stream, &bs_region, &bs_cells, &bs_tok, &bs_pieces, &h_bs};
        for (Buf *b : all) b->release();
    }
};

struct dbg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    static constexpr int MAX_WAVES = 16;
    Slot slot[MAX_WAVES];
    int waves = 16;      // waves the packed host API cuts a large gzip / deflate batch into (cfg2 end to end: 2 -> 23.5, 4 -> 25.8, 8 -> 26.7, 16 -> 27.4 GB/s)
    int png_waves = 8;   // the same for PNG batches (every wave synchronises the host twice on the lane-serial path)
    cudaEvent_t wave_ready = nullptr;
    cudaStream_t up_stream = nullptr;   // packed host API: all uploads, in wave order (one queue, so they finish in that order)
    // dbg_pipe: the packed path calls gate_enter before it enqueues a batch's uploads and gate_leave once they have
    // arrived, so that the batches in flight upload one after the other (and then download one after the other)
    uint32_t in_flight_share = 1;  // dbg_pipe: batches in flight beside this context's
    bool ordered_uploads = false;  // dbg_pipe: gzip / deflate uploads through the one upload stream as well (they arrive in wave order)
    void (*gate_enter)(void *) = nullptr;
    void (*gate_leave)(void *) = nullptr;
    void *gate_arg = nullptr;
    std::atomic<uint64_t> launches{0};
    std::mutex prof_mu;
    char err[512] = "";
    // optional per-launch timing of the dominant (inflate) kernel, for roofline reports
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<int> prof_tag;  // DBG_PROF_* of every bracket
    size_t prof_used = 0;
    // tunables (environment)
    uint32_t fx_expand_ctas = 12;   // CTAs per SM of the lane-serial path's token expansion (DBG_FX_EXPAND_CTAS)
    uint32_t fx_chunk_forced = 0, fx_group_forced = 0;  // DBG_FX_CHUNK / DBG_FX_GROUP: fixed chunk / group size (experiments)
    bool fx = true;                 // lane-serial path for single fixed-Huffman-block streams
    uint32_t bsplit_tok_per_byte = 4;             // token slots per compressed byte (0 = no tokens: decode twice)
    uint64_t bsplit_tok_max_bytes = 24ull << 30;  // the token area never grows beyond this
    bool bs_pieces = true;                // DBG_BS_PIECES: chunks of a handful of long streams are expanded in pieces (cut_token_pieces)
    uint32_t bs_pieces_max_slots = 512;   // ... into at most this many marker domains per call (16 per region when the regions are few)
    bool bsplit = true;
    bool rounds = true;             // the warp-per-stream kernel decodes Huffman blocks in lane-parallel rounds (DBG_ROUNDS)
    uint32_t round_bits = dbg::LB_ROUND_BITS;  // their length (DBG_ROUND_BITS)
    bool bsplit_lanes = true;       // the count pass of the block-split path decodes Huffman blocks lane-parallel
    int bsplit_all = 0;             // 1: every stream of >= bsplit_min_bytes takes the block-split path, not only the batch's
                                    // outliers; 2: every such stream that opens with a dynamic block
    uint64_t bsplit_min_bytes = dbg::BS_MIN_BYTES;
    uint32_t bsplit_factor_q = 8;
    uint32_t bsplit_region = dbg::REGION_BYTES, bsplit_region_min = 16384;
    uint32_t small_batch = 1536;    // packed host API: gzip / deflate batches below this run as one wave with the intra-stream paths on
    bool verify = false;            // opt-in: check gzip CRC32 / ISIZE trailers, zlib Adler-32
    uint32_t inflate_ctas_per_sm = dbg::INFLATE_CTAS_PER_SM;  // resident streams per SM = 4x this (tunable: L2 footprint)
    // counters
    std::atomic<uint64_t> bs_streams{0}, bs_fallbacks{0};
    Buf d_stats;                    // lane-serial path: streams decoded / handed back / extra survivors, counted on the device
    // host-API staging
    Buf d_in, d_out, d_desc;       // arenas + descriptor tables
    Buf h_in, h_out, h_desc;       // pinned mirrors
    dbg_ctx() { h_in.pinned_host = h_out.pinned_host = h_desc.pinned_host = true; }
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

extern "C" int dbg_version(void) { return 200; }

extern "C" int dbg_device_count(void)
{
    // The packed host API overlaps up to 16 waves on streams of their own. Streams share the device's hardware queues
    // (8 by default) and waves on one queue wait for each other: measured 196 ms per cfg2 call with 8 queues, 153 ms
    // with 32. The variable is read when the CUDA context is made, so this only helps when the library is the first
    // CUDA user of the process; otherwise set it in the environment. Never overrides the caller's value.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
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
extern "C" uint64_t dbg_kernel_launches(const dbg_ctx *ctx) { return ctx ? ctx->launches.load() : 0ull; }

extern "C" void dbg_destroy(dbg_ctx *ctx);

// `pre_wave` / `pre_up`: streams made by the caller (dbg_pipe_create makes the wave and upload streams of all its contexts
// back to back, so that they land on different hardware queues); the context owns them from here on.
static dbg_ctx *ctx_create(int device, const cudaStream_t *pre_wave, int n_pre, cudaStream_t pre_up)
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
    // every stream, event and kernel attribute is needed: a half-made context would run waves on the legacy stream
    // The wave streams are created first and back to back: streams are mapped to the device's hardware queues
    // (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default) in creation order, and waves that share a queue wait for each other
    // (measured: with the auxiliary streams created in between, the 16 waves of a cfg2 call ran in two groups of 8,
    // 196 ms per call instead of 155).
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n_pre && i < dbg_ctx::MAX_WAVES; i++) ctx->slot[i].stream = pre_wave[i];
    ctx->up_stream = pre_up;
    for (int i = std::max(0, n_pre); i < dbg_ctx::MAX_WAVES && e == cudaSuccess; i++)
        e = cudaStreamCreateWithFlags(&ctx->slot[i].stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < dbg_ctx::MAX_WAVES && e == cudaSuccess; i++) {
        Slot &sl = ctx->slot[i];
        e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sl.aux_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.aux_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.aux_join, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->wave_ready, cudaEventDisableTiming);
    if (e == cudaSuccess && !ctx->up_stream) e = cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = ctx->d_stats.reserve(128);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_stats.p, 0, 128);
    const size_t smem = sizeof(dbg::InflateSmem) * dbg::INFLATE_WARPS_PER_CTA;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dbg::inflate_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dbg::inflate_batch_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 85);
    if (e != cudaSuccess) {
        set_err(nullptr, "dbg_create: %s", cudaGetErrorString(e));
        dbg_destroy(ctx);
        return nullptr;
    }
    dbg::png_configure_kernels();
    if (const char *v = getenv("DBG_SMALL_BATCH")) ctx->small_batch = (uint32_t)atoi(v);
    if (const char *v = getenv("DBG_FX_CHUNK")) ctx->fx_chunk_forced = (uint32_t)std::min(1 << 20, std::max((int)dbg::FX_MIN_CHUNK, atoi(v))) & ~15u;
    if (const char *v = getenv("DBG_FX_GROUP")) ctx->fx_group_forced = (uint32_t)std::min(1 << 22, std::max(4096, atoi(v))) & ~15u;
    if (const char *v = getenv("DBG_FX")) ctx->fx = atoi(v) != 0;
    if (const char *v = getenv("DBG_FX_EXPAND_CTAS")) ctx->fx_expand_ctas = (uint32_t)std::min(16, std::max(1, atoi(v)));
    if (const char *v = getenv("DBG_WAVES")) ctx->waves = std::min((int)dbg_ctx::MAX_WAVES, std::max(1, atoi(v)));
    if (const char *v = getenv("DBG_PNG_WAVES")) ctx->png_waves = std::min((int)dbg_ctx::MAX_WAVES, std::max(1, atoi(v)));
    if (const char *v = getenv("DBG_BSPLIT")) ctx->bsplit = atoi(v) != 0;
    if (const char *v = getenv("DBG_BSPLIT_LANES")) ctx->bsplit_lanes = atoi(v) != 0;
    if (const char *v = getenv("DBG_BS_PIECES")) ctx->bs_pieces = atoi(v) != 0;
    if (const char *v = getenv("DBG_ROUNDS")) ctx->rounds = atoi(v) != 0;
    if (const char *v = getenv("DBG_ROUND_BITS")) ctx->round_bits = (uint32_t)std::min(1 << 20, std::max(8192, atoi(v)));
    if (const char *v = getenv("DBG_BSPLIT_ALL")) ctx->bsplit_all = atoi(v);
    if (const char *v = getenv("DBG_BSPLIT_FACTOR_Q")) ctx->bsplit_factor_q = (uint32_t)std::max(1, atoi(v));
    if (const char *v = getenv("DBG_BSPLIT_TOKENS")) ctx->bsplit_tok_per_byte = (uint32_t)std::min(8, std::max(0, atoi(v)));
    if (const char *v = getenv("DBG_BSPLIT_REGION")) ctx->bsplit_region = (uint32_t)std::min(1 << 20, std::max(4096, atoi(v)));
    if (const char *v = getenv("DBG_BSPLIT_REGION_MIN")) ctx->bsplit_region_min = (uint32_t)std::min((int)ctx->bsplit_region, std::max(4096, atoi(v)));
    if (const char *v = getenv("DBG_BSPLIT_MIN_BYTES")) ctx->bsplit_min_bytes = std::max<uint64_t>(strtoull(v, nullptr, 10), 2 * (uint64_t)ctx->bsplit_region);
    if (const char *v = getenv("DBG_INFLATE_CTAS_PER_SM")) {
        int k = atoi(v);
        if (k >= 1 && k <= dbg::INFLATE_CTAS_PER_SM) ctx->inflate_ctas_per_sm = (uint32_t)k;
    }
    return ctx;
}

extern "C" dbg_ctx *dbg_create(int device) { return ctx_create(device, nullptr, 0, nullptr); }

extern "C" void dbg_destroy(dbg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    Buf *all[] = {&ctx->d_in, &ctx->d_out, &ctx->d_desc, &ctx->h_in, &ctx->h_out, &ctx->h_desc, &ctx->d_stats};
    for (Buf *b : all) b->release();
    for (int i = 0; i < dbg_ctx::MAX_WAVES; i++) {
        ctx->slot[i].release();
        Slot &sl = ctx->slot[i];
        if (sl.stream) cudaStreamDestroy(sl.stream);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.uploaded) cudaEventDestroy(sl.uploaded);
        if (sl.aux_stream) cudaStreamDestroy(sl.aux_stream);
        if (sl.aux_fork) cudaEventDestroy(sl.aux_fork);
        if (sl.aux_join) cudaEventDestroy(sl.aux_join);
    }
    for (auto &pe : ctx->prof_events) {
        cudaEventDestroy(pe.first);
        cudaEventDestroy(pe.second);
    }
    if (ctx->wave_ready) cudaEventDestroy(ctx->wave_ready);
    if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
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
    // counted on the device (the packed PNG path never waits for them); reading them waits for the device
    uint64_t v[3] = {0, 0, 0};
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(v, ctx->d_stats.p, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess)
        return DBG_ERR_CUDA;
    if (streams) *streams = v[0];
    if (handed_back) *handed_back = v[1];
    if (extra_runs) *extra_runs = v[2];
    return DBG_OK;
}

// Diagnostics of the block-split path's count / expansion passes, counted on the device: v[0] blocks where the
// lane-parallel decode was attempted, v[1] blocks it decoded whole, v[2] blocks it decoded a prefix of, v[3] chunks
// whose tokens were expanded, v[4] chunks that had to be Huffman-decoded a second time. Waits for the device.
extern "C" int dbg_lane_stats(const dbg_ctx *ctx, uint32_t v[8])
{
    if (!ctx || !v) return DBG_ERR_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(v, (uint64_t *)ctx->d_stats.p + 4, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(v + 5, (uint64_t *)ctx->d_stats.p + 8, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
        return DBG_ERR_CUDA;
    return DBG_OK;
}

extern "C" int dbg_set_verify(dbg_ctx *ctx, int on)
{
    if (!ctx) return DBG_ERR_ARG;
    ctx->verify = on != 0;
    return DBG_OK;
}

// Releases the grow-only scratch the context keeps between calls (cells, tokens, PNG scanline buffers, staging
// arenas); the next call allocates what it needs again.
extern "C" int dbg_trim(dbg_ctx *ctx)
{
    if (!ctx) return DBG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < dbg_ctx::MAX_WAVES; i++) ctx->slot[i].release();
    Buf *all[] = {&ctx->d_in, &ctx->d_out, &ctx->d_desc, &ctx->h_in, &ctx->h_out, &ctx->h_desc};
    for (Buf *b : all) b->release();
    return DBG_OK;
}

extern "C" int dbg_profile_enable(dbg_ctx *ctx, int on)
{
    if (!ctx) return DBG_ERR_ARG;
    ctx->profiling = on != 0;
    ctx->prof_used = 0;
    return DBG_OK;
}

// Sum of the device durations (ms) of the brackets with tag `tag` recorded since dbg_profile_enable(ctx, 1), and
// how many there were. Waits for them; dbg_profile_enable(ctx, 1) resets.
extern "C" int dbg_profile_read_tag(dbg_ctx *ctx, int tag, double *total_ms, uint64_t *launches)
{
    if (!ctx || !total_ms || !launches) return DBG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    double sum = 0;
    uint64_t cnt = 0;
    for (size_t i = 0; i < ctx->prof_used; i++) {
        if (ctx->prof_tag[i] != tag) continue;
        float ms = 0;
        CU(cudaEventSynchronize(ctx->prof_events[i].second));
        CU(cudaEventElapsedTime(&ms, ctx->prof_events[i].first, ctx->prof_events[i].second));
        sum += ms;
        cnt++;
    }
    *total_ms = sum;
    *launches = cnt;
    return DBG_OK;
}

// The warp-per-stream inflate kernel's launches (tag DBG_PROF_INFLATE); resets the recording.
extern "C" int dbg_profile_read(dbg_ctx *ctx, double *total_ms, uint64_t *launches)
{
    int rc = dbg_profile_read_tag(ctx, DBG_PROF_INFLATE, total_ms, launches);
    if (rc == DBG_OK) ctx->prof_used = 0;
    return rc;
}

extern "C" int dbg_synchronize(dbg_ctx *ctx)
{
    if (!ctx) return DBG_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return DBG_OK;
}

// ------------------------------------------------------------------ launches --
// Brackets a kernel sequence with the profiling events when dbg_profile_enable() is on.
struct ProfScope {
    dbg_ctx *ctx;
    cudaStream_t s;
    cudaEvent_t e1 = nullptr;
    ProfScope(dbg_ctx *c, cudaStream_t st, int tag) : ctx(c), s(st)
    {
        if (!ctx->profiling) return;
        std::lock_guard<std::mutex> lock(ctx->prof_mu);  // the waves of a PNG batch are driven by threads of their own
        if (ctx->prof_used == ctx->prof_events.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            ctx->prof_events.push_back({a, b});
            ctx->prof_tag.push_back(0);
        }
        ctx->prof_tag[ctx->prof_used] = tag;
        cudaEventRecord(ctx->prof_events[ctx->prof_used].first, s);
        e1 = ctx->prof_events[ctx->prof_used].second;
        ctx->prof_used++;
    }
    ~ProfScope()
    {
        if (e1) cudaEventRecord(e1, s);
    }
};

static int launch_inflate_plain(dbg_ctx *ctx, Slot &sl, dbg::InflateBatch a, int counter_idx, cudaStream_t s, bool may_order = true)
{
    a.counter = (uint32_t *)sl.counter.p + counter_idx;
    if (may_order && !a.order && a.n > (uint32_t)ctx->sm_count) {  // heaviest first, so that the longest streams do not start last
        CU(sl.sched.reserve((size_t)a.n * 4));
        dbg::sched_order_kernel<<<1, 1024, 0, s>>>(a, (uint32_t *)sl.sched.p);
        ctx->launches++;
        a.order = (const uint32_t *)sl.sched.p;
    }
    CU(cudaMemsetAsync(a.counter, 0, sizeof(uint32_t), s));
    uint32_t ctas_needed = (a.n + dbg::INFLATE_WARPS_PER_CTA - 1) / dbg::INFLATE_WARPS_PER_CTA;
    uint32_t grid = std::min<uint32_t>(ctas_needed, (uint32_t)ctx->sm_count * ctx->inflate_ctas_per_sm);
    if (ctx->rounds) {
        // every warp of the grid gets a token scratch of its own. The auxiliary-stream launch of the block-split path
        // and the main launch may run side by side: they use different halves.
        const size_t per_launch = (size_t)ctx->sm_count * ctx->inflate_ctas_per_sm * dbg::INFLATE_WARPS_PER_CTA * dbg::LB_ROUND_TOKENS * 4;
        CU(sl.round_tok.reserve(3 * per_launch));
        a.tok_scratch = (uint32_t *)((uint8_t *)sl.round_tok.p + (size_t)counter_idx * per_launch);
        a.lb_stats = (uint32_t *)((uint64_t *)ctx->d_stats.p + 8);
        a.round_bits = ctx->round_bits;
    }
    size_t smem = sizeof(dbg::InflateSmem) * dbg::INFLATE_WARPS_PER_CTA;
    {
        ProfScope prof(ctx, s, DBG_PROF_INFLATE);
        dbg::inflate_batch_kernel<<<grid, dbg::INFLATE_THREADS, smem, s>>>(a);
    }
    ctx->launches++;
    CU(cudaGetLastError());
    return DBG_OK;
}

// Lane-serial path for streams that are a single fixed-Huffman block (every stb-written PNG): fx_core.h /
// fx_kernels.cuh. Two small device->host reads (how many streams / bytes; exact token and cell counts), so `s`
// is synchronised twice. *skip_out = per-stream flags of the streams handled here, *redo_out = those handed
// back (flag set as well) for a second warp-per-stream pass; *n_redo tells whether there are any.
// What a caller that has the batch on the HOST knows up front (packed API): with it the scratch can be sized by upper
// bounds and nothing is read back from the device, so the call never waits for the stream.
struct FxHints {
    bool have = false;
    uint64_t total_in = 0;     // compressed bytes of the batch (an upper bound of those on this path)
    uint64_t max_in = 0;       // the longest stream
    uint64_t cells_bound = 0;  // sum of the streams' output sizes (upper bound)
};

static int run_fx(dbg_ctx *ctx, Slot &sl, const dbg::InflateBatch &a, cudaStream_t s, const FxHints &hints, const uint32_t **skip_out,
                  const uint32_t **redo_out, uint32_t *n_redo, bool *all_taken)
{
    *skip_out = nullptr;
    *redo_out = nullptr;
    *n_redo = 0;
    *all_taken = false;
    const uint32_t n = a.n;
    CU(sl.h_fx.reserve(sizeof(dbg::FxSummary)));
    CU(sl.fx_stream.reserve(256 + (size_t)n * (2 * 8 + 6 * 4) + 256));
    uint8_t *p = (uint8_t *)sl.fx_stream.p;
    dbg::FxBatch b{};
    b.in_base = a.in_base; b.in_off = a.in_off; b.in_size = a.in_size;
    b.out_base = a.out_base; b.out_off = a.out_off; b.out_cap = a.out_cap;
    b.out_size = a.out_size; b.status = a.status; b.pre_status = a.pre_status; b.n = n;
    b.stats = (uint64_t *)ctx->d_stats.p;
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
    dbg::FxSummary *hs = (dbg::FxSummary *)sl.h_fx.p;
    if (hints.have) {
        hs->n_fx = n;
        hs->fx_in = hints.total_in;
        hs->max_in = hints.max_in;
    } else {
        CU(cudaMemcpyAsync(hs, b.summary, sizeof(dbg::FxSummary), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (hs->n_fx == 0) return DBG_OK;
    }
    // chunk = what one lane decodes: 16 KiB, halved (down to 2 KiB) while the batch has fewer chunks than the
    // GPU has lanes to give; group = what one warp expands and the unit of the marker / resolve scheme:
    // 256 KiB of compressed data, smaller for small batches (more warps), larger for very long streams (the
    // last 32 KiB of every group's output is resolved by a per-stream serial chain).
    const uint64_t lanes_wanted = (uint64_t)ctx->sm_count * 2048;
    uint32_t chunk = 16384;
    while (chunk > dbg::FX_MIN_CHUNK && hs->fx_in / chunk < lanes_wanted) chunk >>= 1;
    if (ctx->fx_chunk_forced) chunk = ctx->fx_chunk_forced;
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
    CU(sl.fx_chunks.reserve(per_chunk));
    uint8_t *q = (uint8_t *)sl.fx_chunks.p;
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
    CU(cudaMemsetAsync(b.chunk_stream, 0, (size_t)T * 2 * 4, s));     // chunk_stream, nsurv of unused slots
    CU(cudaMemsetAsync(b.c_surv, 0xff, (size_t)T * 4, s));            // FX_NONE
    CU(cudaMemsetAsync(b.group_stream, 0, (size_t)NG * 4 * 4, s));    // group_stream, g_out_len, g_flag, g_ntok of unused slots
    const uint32_t warp_grid = std::min<uint32_t>((T + dbg::FX_WARPS_PER_CTA - 1) / dbg::FX_WARPS_PER_CTA, (uint32_t)ctx->sm_count * 12);
    const uint32_t lane_grid = std::min<uint32_t>((T + dbg::FX_LANE_THREADS - 1) / dbg::FX_LANE_THREADS, (uint32_t)ctx->sm_count * 12);
    {
        ProfScope prof(ctx, s, DBG_PROF_FX_SIZES);
        dbg::fx_assign_kernel<<<sb, 128, 0, s>>>(b);
        dbg::fx_fill_kernel<<<n, 128, 0, s>>>(b);
        dbg::fx_head_kernel<<<warp_grid, dbg::FX_WARPS_PER_CTA * 32, 0, s>>>(b);
        dbg::fx_sizes_kernel<<<lane_grid, dbg::FX_LANE_THREADS, 0, s>>>(b);
    }
    ctx->launches += 4;
    CU(cudaGetLastError());
    // the chain kernel hands a stream back when the cells / tokens do not fit: exact sizes after a read-back, else
    // upper bounds (a token takes at least 8 bits of a fixed-Huffman stream)
    if (hints.have) {
        b.cells_cap = hints.cells_bound;
        b.tok_cap = hs->fx_in + n;
        CU(sl.cells.reserve((size_t)b.cells_cap * 2 + 256));
        CU(sl.fx_tok.reserve((size_t)b.tok_cap * 4 + 256));
    } else {
        b.cells_cap = b.tok_cap = ~0ull;
    }
    dbg::fx_chain_kernel<<<std::min<uint32_t>((n + dbg::FX_WARPS_PER_CTA - 1) / dbg::FX_WARPS_PER_CTA, (uint32_t)ctx->sm_count * 8),
                           dbg::FX_WARPS_PER_CTA * 32, 0, s>>>(b);
    ctx->launches++;
    CU(cudaGetLastError());
    if (!hints.have) {
        CU(cudaMemcpyAsync(hs, b.summary, sizeof(dbg::FxSummary), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    if (hints.have || hs->cells_used) {
        if (!hints.have) {
            CU(sl.cells.reserve((size_t)hs->cells_used * 2 + 256));
            CU(sl.fx_tok.reserve((size_t)hs->tok_used * 4 + 256));
        }
        b.cells = (uint16_t *)sl.cells.p;
        b.tok = (uint32_t *)sl.fx_tok.p;
        ProfScope prof(ctx, s, DBG_PROF_FX_EXPAND);
        dbg::fx_tokens_kernel<<<lane_grid, dbg::FX_LANE_THREADS, 0, s>>>(b);
        dbg::fx_expand_kernel<<<std::min<uint32_t>((NG + dbg::FX_WARPS_PER_CTA - 1) / dbg::FX_WARPS_PER_CTA, (uint32_t)ctx->sm_count * ctx->fx_expand_ctas),
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
    *n_redo = hints.have ? 1u : hs->n_redo;  // not known without the read-back: the second pass is launched, and finds nothing to do
    *all_taken = !hints.have && hs->n_fx == n && hs->n_redo == 0;
    return DBG_OK;
}

// Joins the auxiliary stream into `s` when run_bsplit leaves early: the kernel forked onto it reads the caller's
// buffers and the slot's work queue, so nothing the caller enqueues next may overtake it.
struct AuxJoin {
    Slot &sl;
    cudaStream_t s;
    bool armed = false;
    ~AuxJoin()
    {
        if (armed) cudaStreamWaitEvent(s, sl.aux_join, 0);
    }
};

// Block-split path (bsplit_kernels.cuh): long multi-block streams. When streams do take this path, the warp-per-stream
// kernel for all the others is launched from here, on an auxiliary stream, as soon as the classification is known: it
// is bound by the latency of its longest streams and leaves most SM slots free, which the (throughput-bound)
// block-split kernels then fill. *regular_done tells the caller that this has happened.
static int run_bsplit(dbg_ctx *ctx, Slot &sl, dbg::InflateBatch a, cudaStream_t s, const uint32_t *taken, bool *regular_done)
{
    *regular_done = false;
    const uint32_t n = a.n;
    CU(sl.h_bs.reserve(sizeof(dbg::BsSummary)));
    CU(sl.bs_stream.reserve(256 + (size_t)n * (8 + 8 + 4 + 4 + 4 + 4) + 256));
    uint8_t *p = (uint8_t *)sl.bs_stream.p;
    dbg::BsBatch b{};
    b.in_base = a.in_base; b.in_off = a.in_off; b.in_size = a.in_size;
    b.out_base = a.out_base; b.out_off = a.out_off; b.out_cap = a.out_cap;
    b.out_size = a.out_size; b.status = a.status; b.pre_status = a.pre_status; b.taken = taken; b.n = n;
    b.resident_warps = (uint32_t)ctx->sm_count * ctx->inflate_ctas_per_sm * dbg::INFLATE_WARPS_PER_CTA;
    b.min_bytes = ctx->bsplit_min_bytes;
    b.lanes = (ctx->bsplit_lanes && ctx->bsplit_tok_per_byte) ? 1u : 0u;
    b.lb_stats = (uint32_t *)((uint64_t *)ctx->d_stats.p + 4);
    // which streams: those that would keep one warp busy clearly longer than the batch's fair share (factor_q / 4 x batch
    // bytes / resident warps). With lane-parallel blocks EVERY stream of >= min_bytes could take this path (DBG_BSPLIT_ALL=1),
    // but measured on cfg2 that loses: search + count + expansion + resolve cost ~40 warp instructions per symbol, against
    // ~64 for the plain warp-per-stream kernel that needs no second pass, no cells and no tokens
    b.factor_q = ctx->bsplit_all == 1 ? 0u : ctx->bsplit_factor_q;
    b.dyn_all = (ctx->bsplit_all == 2 && b.lanes) ? 1u : 0u;
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
    dbg::BsSummary *hs = (dbg::BsSummary *)sl.h_bs.p;
    CU(cudaMemcpyAsync(hs, b.summary, sizeof(dbg::BsSummary), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (hs->n_split == 0) return DBG_OK;
    // fork: everything that is not split, on the auxiliary stream
    a.skip = taken;
    a.skip2 = b.flag;
    AuxJoin join{sl, s};
    CU(cudaEventRecord(sl.aux_fork, s));
    CU(cudaStreamWaitEvent(sl.aux_stream, sl.aux_fork, 0));
    {
        int rc = launch_inflate_plain(ctx, sl, a, 0, sl.aux_stream);
        cudaError_t er = cudaEventRecord(sl.aux_join, sl.aux_stream);
        join.armed = er == cudaSuccess;
        if (rc) return rc;
        CU(er);
    }
    *regular_done = true;
    // region size: 64 KiB when that already gives every resident warp a few regions, else smaller (more, shorter
    // chunks: the latency of a lone long stream is the decode time of its longest chunk)
    uint32_t region = ctx->bsplit_region;
    while (region > ctx->bsplit_region_min && hs->split_in / region < 4ull * b.resident_warps) region >>= 1;
    b.region_bytes = region;
    const uint32_t T = (uint32_t)(hs->split_in / region) + hs->n_split;  // upper bound of the region count
    CU(sl.bs_region.reserve((size_t)T * (4 + 8 + 8 + 8 + 4 + 4 + 4) + 256));
    CU(cudaMemsetAsync(sl.bs_region.p, 0, (size_t)T * (4 + 8 + 8 + 8 + 4 + 4 + 4), s));
    b.cand = (uint64_t *)sl.bs_region.p;
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
        if (sl.bs_tok.reserve((size_t)tpb * 4 * hs->split_in + 256) == cudaSuccess) {
            b.tok = (uint32_t *)sl.bs_tok.p;
            b.tok_per_byte = tpb;
        } else {
            (void)cudaGetLastError();  // no room for tokens: the second pass decodes the Huffman codes again
        }
    }
    if (!b.tok) b.lanes = 0;
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
        CU(sl.bs_cells.reserve((size_t)hs->cells_used * 2 + 256));
        b.cells = (uint16_t *)sl.bs_cells.p;
        dbg::SplitBatch r{};
        r.out_base = a.out_base; r.out_off = a.out_off; r.out_size = a.out_size; r.status = a.status; r.n = n;
        r.split_flag = b.flag; r.redo = b.redo; r.cell_base = b.cell_base; r.cells = b.cells;
        // A handful of long streams (few regions): the expansion of a chunk is one warp's latency chain, so every chunk's token
        // run is cut into up to 16 pieces, each expanded by a warp of its own and resolved as a marker domain of its own
        // (gzipsample.gz as a batch of one: 2.3 of its 3.9 ms were that chain). Large batches have chunks enough.
        // (at most bs_pieces_max_slots domains in all: the tails of a stream's domains are resolved one after the other)
        const uint32_t M = (ctx->bs_pieces && b.tok) ? std::min<uint32_t>(16u, std::max<uint32_t>(1u, ctx->bs_pieces_max_slots / std::max<uint32_t>(T, 1u))) : 1u;
        uint32_t domains = T;
        if (M > 1) {
            const size_t slots = (size_t)T * M;
            CU(sl.bs_pieces.reserve(slots * (8 + 4 * 4) + (size_t)n * 8 + 256));
            CU(cudaMemsetAsync(sl.bs_pieces.p, 0, slots * (8 + 4 * 4) + (size_t)n * 8, s));
            dbg::BsPieces q{};
            q.max_pieces = M;
            q.p_out_off = (uint64_t *)sl.bs_pieces.p;
            q.p_stream = (uint32_t *)(q.p_out_off + slots);
            q.p_out_len = q.p_stream + slots;
            q.p_flag = q.p_out_len + slots;
            q.p_ntok = q.p_flag + slots;
            q.p_base = q.p_ntok + slots;
            q.p_count = q.p_base + n;
            dbg::bs_piece_ranges_kernel<<<sb, 128, 0, s>>>(b, q);
            dbg::bs_pieces_kernel<<<grid, dbg::BS_WARPS_PER_CTA * 32, 0, s>>>(b, q);
            const uint32_t pgrid = std::min<uint32_t>((uint32_t)((slots + dbg::BS_WARPS_PER_CTA - 1) / dbg::BS_WARPS_PER_CTA),
                                                      (uint32_t)ctx->sm_count * dbg::INFLATE_CTAS_PER_SM);
            dbg::bs_decode_pieces_kernel<<<pgrid, dbg::BS_WARPS_PER_CTA * 32, smem, s>>>(b, q);
            ctx->launches += 2;
            r.chunk_base = q.p_base; r.nchunks = q.p_count; r.chunk_stream = q.p_stream; r.c_out_off = q.p_out_off;
            r.c_out_len = q.p_out_len; r.c_flag = q.p_flag;
            domains = (uint32_t)slots;
        } else {
            dbg::bs_decode_kernel<<<grid, dbg::BS_WARPS_PER_CTA * 32, smem, s>>>(b);
            r.chunk_base = b.chunk_base; r.nchunks = b.nchunks; r.chunk_stream = b.chunk_stream; r.c_out_off = b.c_out_off;
            r.c_out_len = b.c_out_len; r.c_flag = b.c_flag;
        }
        // cells -> bytes with the resolve kernels
        dbg::split_resolve_tails_kernel<<<n, dbg::RESOLVE_THREADS, 0, s>>>(r);
        dbg::split_resolve_body_kernel<<<std::min<uint32_t>(domains, (uint32_t)ctx->sm_count * 8), 256, 0, s>>>(r, domains);
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
        int rc = launch_inflate_plain(ctx, sl, again, 1, s, false);
        if (rc) return rc;
    }
    return DBG_OK;  // ~AuxJoin makes `s` wait for the auxiliary stream
}

// All decode paths for one batch of raw DEFLATE streams. `intra` = the paths that cut long streams into pieces may
// be used (they synchronise the host with `s`).
static int launch_inflate(dbg_ctx *ctx, Slot &sl, dbg::InflateBatch a, cudaStream_t s, bool intra_fx, bool intra_bs,
                          const FxHints &hints = FxHints())
{
    CU(sl.counter.reserve(8 * sizeof(uint32_t)));
    a.skip = nullptr;
    a.skip2 = nullptr;
    const uint32_t *fx_redo = nullptr;
    uint32_t n_redo = 0;
    if (ctx->fx && intra_fx) {
        const uint32_t *skip = nullptr;
        bool all_taken = false;
        int rc = run_fx(ctx, sl, a, s, hints, &skip, &fx_redo, &n_redo, &all_taken);
        if (rc) return rc;
        a.skip = skip;
        if (all_taken) return DBG_OK;  // (a batch of stb-written PNGs) nothing is left for the other paths
    }
    if (n_redo) {
        // streams the lane-serial path handed back: a warp-per-stream pass of their own (a.skip keeps them out of
        // everything below)
        dbg::InflateBatch again = a;
        again.skip = nullptr;
        again.only = fx_redo;
        again.order = nullptr;
        int rc = launch_inflate_plain(ctx, sl, again, 2, s, false);
        if (rc) return rc;
    }
    if (ctx->bsplit && intra_bs && !hints.have) {  // (with hints the caller has seen that every stream is a fixed block)
        bool done = false;
        int rc = run_bsplit(ctx, sl, a, s, a.skip, &done);
        if (rc || done) return rc;
    }
    return launch_inflate_plain(ctx, sl, a, 0, s);
}

// Raw deflate (gz == false) or gzip members (gz == true) of one slot.
static int inflate_device_slot(dbg_ctx *ctx, Slot &sl, bool gz, bool intra, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                               const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                               const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_status,
                               const uint32_t *d_order, cudaStream_t s)
{
    if (gz) {
        CU(sl.meta.reserve(n * 20 + 64));
        uint64_t *gz_off = (uint64_t *)sl.meta.p;
        uint64_t *gz_size = gz_off + n;
        uint32_t *gz_pre = (uint32_t *)(gz_size + n);
        dbg::gz_scan_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_in, d_in_off, d_in_size, (uint32_t)n, gz_off, gz_size, gz_pre);
        ctx->launches++;
        CU(cudaGetLastError());
        dbg::InflateBatch a{d_in, gz_off, gz_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, gz_pre, nullptr, nullptr, nullptr, d_order, nullptr, (uint32_t)n, nullptr, nullptr, 0};
        int rc = launch_inflate(ctx, sl, a, s, intra, intra);
        if (rc || !ctx->verify) return rc;
        uint32_t ctas = (uint32_t)std::min<uint64_t>((n + dbg::SCAN_WARPS - 1) / dbg::SCAN_WARPS, (uint64_t)ctx->sm_count * 8);
        dbg::gz_verify_kernel<<<ctas, dbg::SCAN_WARPS * 32, 0, s>>>(d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_size, d_status,
                                                                    (uint32_t)n);
        ctx->launches++;
        CU(cudaGetLastError());
        return DBG_OK;
    }
    dbg::InflateBatch a{d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, nullptr, nullptr, nullptr, nullptr, d_order, nullptr, (uint32_t)n, nullptr, nullptr, 0};
    return launch_inflate(ctx, sl, a, s, intra, intra);
}

static int png_device_slot(dbg_ctx *ctx, Slot &sl, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                           const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                           uint32_t *d_status, uint64_t total_in_bytes, uint64_t total_rgba_bytes, cudaStream_t s,
                           const FxHints &hints = FxHints())
{
    CU(sl.png_scratch.reserve(dbg::png_scratch_bytes(n, total_in_bytes, total_rgba_bytes)));
    dbg::PngLayout lay = dbg::png_layout((uint8_t *)sl.png_scratch.p, n, total_in_bytes, total_rgba_bytes);
    // 1. container walk + CRC-32 + IDAT gather (one warp per image)
    dbg::PngBatch pb{d_in, d_in_off, d_in_size, d_out_cap, (uint32_t)n, lay};
    int rc;
    {
        ProfScope prof(ctx, s, DBG_PROF_PNG_SCAN);
        rc = dbg::png_launch_scan(pb, ctx->sm_count, s);
    }
    ctx->launches += 4;
    if (rc) CU((cudaError_t)rc);
    // 2. inflate the zlib payloads into the filtered-scanline buffers
    // z_off holds absolute addresses: a single-IDAT image is inflated straight from the file
    dbg::InflateBatch a{nullptr, lay.z_off, lay.z_size, lay.scan, lay.s_off, lay.s_cap, lay.s_size, lay.inf_status,
                        lay.pre_status, nullptr, nullptr, nullptr, nullptr, nullptr, (uint32_t)n, nullptr, nullptr, 0};
    rc = launch_inflate(ctx, sl, a, s, true, true, hints);
    if (rc) return rc;
    if (ctx->verify) {
        uint32_t ctas = (uint32_t)std::min<uint64_t>((n + dbg::SCAN_WARPS - 1) / dbg::SCAN_WARPS, (uint64_t)ctx->sm_count * 8);
        dbg::png_adler_kernel<<<ctas, dbg::SCAN_WARPS * 32, 0, s>>>(pb);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    // 3. un-filter (+ palette / RGB expansion) straight into the caller's RGBA
    {
        ProfScope prof(ctx, s, DBG_PROF_PNG_UNFILTER);
        rc = dbg::png_launch_unfilter(pb, d_out, d_out_off, d_status, ctx->sm_count, s);
    }
    ctx->launches += 2;
    if (rc) CU((cudaError_t)rc);
    return DBG_OK;
}

// The device-resident entry points share slot 0. A call on another stream than the previous one must not touch that
// scratch while the previous call's kernels still use it: every call waits (on the device) for the event the
// previous one left behind, and leaves its own -- also when it fails half way.
struct SlotGuard {
    Slot &sl;
    cudaStream_t s;
    SlotGuard(Slot &slot, cudaStream_t st) : sl(slot), s(st) { cudaStreamWaitEvent(s, sl.done, 0); }
    ~SlotGuard() { cudaEventRecord(sl.done, s); }
};

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
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    SlotGuard guard(ctx->slot[0], s);
    return inflate_device_slot(ctx, ctx->slot[0], false, true, n, d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size,
                               d_status, d_order, s);
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
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    SlotGuard guard(ctx->slot[0], s);
    return inflate_device_slot(ctx, ctx->slot[0], true, true, n, d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size,
                               d_status, d_order, s);
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
    SlotGuard guard(ctx->slot[0], s);
    return png_device_slot(ctx, ctx->slot[0], n, d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_status, total_in_bytes,
                           total_rgba_bytes, s);
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

// Host-side look at a PNG batch: true when every file has a plausible IHDR and its first IDAT opens with a zlib header
// followed by a FINAL FIXED-Huffman block (what stb_image_write emits, stb_write.h:913-916). est[i] = w*h*4 + h + 1, the
// reference's bound of the filtered scanlines (decode_png.c:965-968). Nothing is decided here about validity: the
// device-side chunk walk does that; a file that merely looks right costs nothing but scratch.
static bool png_all_stb_shaped(uint64_t n, const uint8_t *h_in, const uint64_t *in_off, const uint64_t *in_size, std::vector<uint64_t> &est)
{
    est.resize(n);
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *f = h_in + in_off[i];
        const uint64_t size = in_size[i];
        if (size < 8 + 25 + 12 + 3) return false;
        auto be = [&](uint64_t at) { return ((uint64_t)f[at] << 24) | ((uint64_t)f[at + 1] << 16) | ((uint64_t)f[at + 2] << 8) | f[at + 3]; };
        const uint64_t w = be(16), h = be(20);
        if (w == 0 || h == 0 || w * h * 4 + h + 1 >= (1ull << 32)) return false;
        est[i] = w * h * 4 + h + 1;
        uint64_t pos = 8;
        bool found = false;
        for (int c = 0; c < 64 && pos + 12 <= size; c++) {
            const uint64_t len = be(pos);
            if (f[pos + 4] == 'I' && f[pos + 5] == 'D' && f[pos + 6] == 'A' && f[pos + 7] == 'T') {
                found = len >= 3 && pos + 11 <= size && (f[pos + 10] & 7) == 3;
                break;
            }
            pos += 12 + len;
        }
        if (!found) return false;
    }
    return true;
}

// The packed path proper. `cuts` are the wave boundaries (item indices, cuts.front() == 0, cuts.back() == n); the
// items of a wave must lie in index order, without overlap, in both host arenas (one upload and one download per
// wave). The waves are packed next to each other in the context's device arenas, wherever they lie on the host, so
// a caller may pass any sub-set of a large arena (the multi-device entry point does). `intra` = the intra-stream
// paths may be used (gzip / deflate; PNG waves always use them).
static int packed_waves(dbg_ctx *ctx, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off, const uint64_t *in_size,
                        uint8_t *h_out, const uint64_t *out_off, const uint64_t *out_cap, uint64_t *out_size, uint32_t *status,
                        const std::vector<uint64_t> &cuts, bool intra)
{
    // PNG: per-image size of the filtered scanlines when every file of the batch is stb-shaped (else empty)
    std::vector<uint64_t> png_est;
    if (kind == 2 && png_all_stb_shaped(n, h_in, in_off, in_size, png_est) == false) png_est.clear();
    const uint64_t *png_hints = png_est.empty() ? nullptr : png_est.data();
    const int nw = (int)cuts.size() - 1;
    cudaStream_t s = ctx->stream;
    // descriptor block: in_off, in_size, out_off, out_cap, out_size (u64 x n each), status + order (u32 x n each)
    const size_t desc_bytes = n * (5 * 8 + 2 * 4);
    CU(ctx->h_desc.reserve(desc_bytes));
    CU(ctx->d_desc.reserve(desc_bytes));
    uint64_t *hd = (uint64_t *)ctx->h_desc.p;
    uint32_t *h_status = (uint32_t *)(hd + 5 * n);
    uint32_t *h_order = h_status + n;
    uint64_t *dd = (uint64_t *)ctx->d_desc.p;
    uint32_t *d_status = (uint32_t *)(dd + 5 * n);
    uint32_t *d_order = d_status + n;
    // device placement of the waves + device-side offsets of the items
    struct Wave { uint64_t b, e, hi0, hi1, ho0, ho1, di, dout, tot_in, tot_out; };
    std::vector<Wave> wv(nw);
    uint64_t din = 0, dout = 0;
    for (int k = 0; k < nw; k++) {
        Wave &w = wv[k];
        w.b = cuts[k];
        w.e = cuts[k + 1];
        w.hi0 = in_off[w.b] & ~(uint64_t)15;  // keep every item's address modulo 16
        w.hi1 = in_off[w.e - 1] + in_size[w.e - 1];
        w.ho0 = out_off[w.b] & ~(uint64_t)15;
        w.ho1 = out_off[w.e - 1] + out_cap[w.e - 1];
        w.di = din;
        w.dout = dout;
        w.tot_in = w.tot_out = 0;
        din += align_up(w.hi1 - w.hi0 + 64, 256);
        dout += align_up(w.ho1 - w.ho0 + 64, 256);
        for (uint64_t i = w.b; i < w.e; i++) {
            hd[i] = in_off[i] - w.hi0 + w.di;
            hd[n + i] = in_size[i];
            hd[2 * n + i] = out_off[i] - w.ho0 + w.dout;
            hd[3 * n + i] = out_cap[i];
            w.tot_in += in_size[i];
            w.tot_out += out_cap[i];
        }
        // per-wave largest-first order (indices relative to the wave)
        uint32_t *o = h_order + w.b;
        const uint64_t m = w.e - w.b;
        std::vector<uint64_t> wt(m);
        for (uint64_t i = 0; i < m; i++) wt[i] = sched_weight(kind, h_in + in_off[w.b + i], in_size[w.b + i]) + out_cap[w.b + i] / 32;
        std::iota(o, o + m, 0u);
        std::stable_sort(o, o + m, [&](uint32_t x, uint32_t y) { return wt[x] > wt[y]; });
    }
    CU(ctx->d_in.reserve(din + 64));
    CU(ctx->d_out.reserve(dout + 64));
    uint8_t *d_in = (uint8_t *)ctx->d_in.p, *d_out = (uint8_t *)ctx->d_out.p;
    struct Gate {  // left on every path out of this function
        dbg_ctx *c;
        bool in = false;
        void enter()
        {
            if (c->gate_enter) c->gate_enter(c->gate_arg);
            in = true;
        }
        void leave()
        {
            if (in && c->gate_leave) c->gate_leave(c->gate_arg);
            in = false;
        }
        ~Gate() { leave(); }
    } gate{ctx};
    gate.enter();
    CU(cudaMemcpyAsync(dd, hd, 4 * n * 8, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_order, h_order, n * 4, cudaMemcpyHostToDevice, s));
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
    // The uploads all go through ONE stream, in wave order: spread over the waves' streams they complete in whatever
    // order the copy queues pick (measured with DBG_WAVE_TRACE: 0, 4, 1, 5, ...), and a wave that is processed early
    // would wait for an upload that was queued late.
    int rc = DBG_OK;
    // PNG: the uploads all go through ONE stream (see above); gzip / deflate: through the waves' own streams (measured:
    // 155 ms per cfg2 call against 193 ms through one upload stream -- the kernels of 16 waves then start in two groups)
    const bool one_up = kind == 2 || ctx->ordered_uploads;
    if (one_up && cudaStreamWaitEvent(ctx->up_stream, ctx->wave_ready, 0) != cudaSuccess) rc = DBG_ERR_CUDA;
    for (int k = 0; k < nw && rc == DBG_OK; k++) {
        const Wave &w = wv[k];
        Slot &sl = ctx->slot[k];
        cudaStream_t ws = sl.stream, us = one_up ? ctx->up_stream : ws;
        cudaError_t e = one_up ? cudaSuccess : cudaStreamWaitEvent(ws, ctx->wave_ready, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_in + w.di, h_in + w.hi0, w.hi1 - w.hi0, cudaMemcpyHostToDevice, us);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_in + w.di + (w.hi1 - w.hi0), 0, 64, us);  // what readers see behind the last item
        if (e == cudaSuccess) e = cudaEventRecord(sl.uploaded, us);
        if (e == cudaSuccess && one_up) e = cudaStreamWaitEvent(ws, sl.uploaded, 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ws, sl.done, 0);  // a device-resident call may still be using this slot's scratch
        if (e != cudaSuccess) {
            set_err(ctx, "packed batch, upload of wave %d: %s", k, cudaGetErrorString(e));
            rc = DBG_ERR_CUDA;
        }
        if (trace) cudaEventRecord(tev[3 * k], ws);
    }
    for (int k = 0; k < nw && rc == DBG_OK && kind == 2; k++) {
        // PNG: when the host has seen that every file is what stb writes (one IDAT, one final fixed-Huffman block),
        // the lane-serial path gets its sizes as hints and the wave is enqueued without ever waiting for the device:
        // the waves then run first in, first out, and the download of one overlaps the kernels of the next. A mixed
        // batch takes the read-back route (two host waits per wave).
        const Wave &w = wv[k];
        Slot &sl = ctx->slot[k];
        const uint64_t b = w.b, m = w.e - w.b;
        FxHints hints;
        hints.have = png_hints != nullptr;
        if (hints.have) {
            hints.total_in = w.tot_in;
            for (uint64_t i = w.b; i < w.e; i++) {
                hints.max_in = std::max(hints.max_in, in_size[i]);
                hints.cells_bound += png_hints[i];
            }
        }
        rc = png_device_slot(ctx, sl, m, d_in, dd + b, dd + n + b, d_out, dd + 2 * n + b, dd + 3 * n + b, d_status + b, w.tot_in,
                             w.tot_out, sl.stream, hints);
        if (trace) cudaEventRecord(tev[3 * k + 1], sl.stream);
        if (rc == DBG_OK && cudaMemcpyAsync(h_out + w.ho0, d_out + w.dout, w.ho1 - w.ho0, cudaMemcpyDeviceToHost, sl.stream) != cudaSuccess)
            rc = DBG_ERR_CUDA;
        if (rc == DBG_OK && cudaMemcpyAsync(h_status + b, d_status + b, m * 4, cudaMemcpyDeviceToHost, sl.stream) != cudaSuccess) rc = DBG_ERR_CUDA;
        if (trace) cudaEventRecord(tev[3 * k + 2], sl.stream);
    }
    for (int k = 0; k < nw && rc == DBG_OK && kind != 2; k++) {
        const Wave &w = wv[k];
        Slot &sl = ctx->slot[k];
        const uint64_t b = w.b, m = w.e - w.b;
        rc = inflate_device_slot(ctx, sl, kind == 1, intra, m, d_in, dd + b, dd + n + b, d_out, dd + 2 * n + b, dd + 3 * n + b,
                                 dd + 4 * n + b, d_status + b, d_order + b, sl.stream);
        if (trace) cudaEventRecord(tev[3 * k + 1], sl.stream);
    }
    if (kind != 2)
        for (int k = 0; k < nw && rc == DBG_OK; k++) {
            const Wave &w = wv[k];
            if (cudaMemcpyAsync(h_out + w.ho0, d_out + w.dout, w.ho1 - w.ho0, cudaMemcpyDeviceToHost, ctx->slot[k].stream) != cudaSuccess)
                rc = DBG_ERR_CUDA;
            // the wave's sizes and statuses travel behind its payload, on its stream: a copy queued on another stream at the
            // end of the call would wait behind every download that another batch in flight (dbg_pipe) has queued already
            const uint64_t b = w.b, m = w.e - w.b;
            if (rc == DBG_OK && (cudaMemcpyAsync(hd + 4 * n + b, dd + 4 * n + b, m * 8, cudaMemcpyDeviceToHost, ctx->slot[k].stream) != cudaSuccess ||
                                 cudaMemcpyAsync(h_status + b, d_status + b, m * 4, cudaMemcpyDeviceToHost, ctx->slot[k].stream) != cudaSuccess))
                rc = DBG_ERR_CUDA;
            if (trace) cudaEventRecord(tev[3 * k + 2], ctx->slot[k].stream);
        }
    if (ctx->gate_leave) {  // the next batch in flight may upload once this one's input has arrived
        for (int k = 0; k < nw && rc == DBG_OK; k++) cudaEventSynchronize(ctx->slot[k].uploaded);
        gate.leave();
    }
    // also on failure: nothing of this call may still be running when it returns
    for (int k = 0; k < nw; k++) {
        cudaEventRecord(ctx->slot[k].done, ctx->slot[k].stream);
        if (cudaStreamSynchronize(ctx->slot[k].stream) != cudaSuccess && rc == DBG_OK) rc = DBG_ERR_CUDA;
    }
    if (rc == DBG_ERR_CUDA && !ctx->err[0]) set_err(ctx, "packed batch: %s", cudaGetErrorString(cudaGetLastError()));
    if (trace) {
        static std::mutex trace_mu;
        static cudaEvent_t origin = nullptr;  // the first traced call's start: calls of several contexts on one time axis
        std::lock_guard<std::mutex> tl(trace_mu);
        float t_call = 0;
        if (!origin) {
            origin = tev[3 * nw];
            tev[3 * nw] = nullptr;
            cudaEventCreate(&tev[3 * nw]);
            cudaEventRecord(tev[3 * nw], s);  // placeholder so that the destroy loop below stays simple
            cudaEventSynchronize(tev[3 * nw]);
        } else {
            cudaEventElapsedTime(&t_call, origin, tev[3 * nw]);
        }
        fprintf(stderr, "ctx %p: call began at %.2f ms\n", (void *)ctx, t_call);
        cudaEvent_t t0ev = t_call == 0 && tev[3 * nw] != origin ? origin : tev[3 * nw];
        for (int k = 0; k < nw; k++) {
            float a = 0, b2 = 0, c = 0;
            cudaEventElapsedTime(&a, t0ev, tev[3 * k]);
            cudaEventElapsedTime(&b2, t0ev, tev[3 * k + 1]);
            cudaEventElapsedTime(&c, t0ev, tev[3 * k + 2]);
            fprintf(stderr, "wave %2d: h2d done %7.2f ms, kernels done %7.2f ms, d2h done %7.2f ms\n", k, a, b2, c);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    if (rc) return rc;
    CU(cudaStreamSynchronize(s));
    memcpy(status, h_status, n * 4);
    if (kind == 2) {
        for (uint64_t i = 0; i < n; i++) out_size[i] = status[i] == 0 ? out_cap[i] : 0;
    } else {
        memcpy(out_size, hd + 4 * n, n * 8);
    }
    return DBG_OK;
}

// One item, or a handful: the download waits for the sizes and moves only what was produced (the reference's
// decode_gz() sizes its buffer at 35 x the input, decode_gz.c:245-247; copying that back would dominate the call).
static int packed_small(dbg_ctx *ctx, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off, const uint64_t *in_size,
                        uint8_t *h_out, const uint64_t *out_off, const uint64_t *out_cap, uint64_t *out_size, uint32_t *status)
{
    cudaStream_t s = ctx->stream;
    Slot &sl = ctx->slot[0];
    const size_t desc_bytes = n * (5 * 8 + 2 * 4);
    CU(ctx->h_desc.reserve(desc_bytes));
    CU(ctx->d_desc.reserve(desc_bytes));
    uint64_t *hd = (uint64_t *)ctx->h_desc.p, *dd = (uint64_t *)ctx->d_desc.p;
    uint32_t *h_status = (uint32_t *)(hd + 5 * n), *d_status = (uint32_t *)(dd + 5 * n);
    uint64_t din = 0, dout = 0, tot_in = 0, tot_out = 0;
    for (uint64_t i = 0; i < n; i++) {
        hd[i] = din + (in_off[i] & 15);
        hd[n + i] = in_size[i];
        hd[2 * n + i] = dout;
        hd[3 * n + i] = out_cap[i];
        din += align_up((in_off[i] & 15) + in_size[i] + 64, 256);
        dout += align_up(out_cap[i] + 64, 256);
        tot_in += in_size[i];
        tot_out += out_cap[i];
    }
    CU(ctx->d_in.reserve(din + 64));
    CU(ctx->d_out.reserve(dout + 64));
    uint8_t *d_in = (uint8_t *)ctx->d_in.p, *d_out = (uint8_t *)ctx->d_out.p;
    SlotGuard guard(sl, s);
    CU(cudaMemcpyAsync(dd, hd, 4 * n * 8, cudaMemcpyHostToDevice, s));
    for (uint64_t i = 0; i < n; i++) {
        CU(cudaMemcpyAsync(d_in + hd[i], h_in + in_off[i], in_size[i], cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync(d_in + hd[i] + in_size[i], 0, 64, s));
    }
    int rc;
    if (kind == 2)
        rc = png_device_slot(ctx, sl, n, d_in, dd, dd + n, d_out, dd + 2 * n, dd + 3 * n, d_status, tot_in, tot_out, s);
    else
        rc = inflate_device_slot(ctx, sl, kind == 1, true, n, d_in, dd, dd + n, d_out, dd + 2 * n, dd + 3 * n, dd + 4 * n, d_status,
                                 nullptr, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(hd + 4 * n, dd + 4 * n, n * 8 + n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (uint64_t i = 0; i < n; i++) {
        status[i] = h_status[i];
        out_size[i] = kind == 2 ? (status[i] == 0 ? out_cap[i] : 0) : hd[4 * n + i];
        if (status[i] == 0 && out_size[i])
            CU(cudaMemcpyAsync(h_out + out_off[i], d_out + hd[2 * n + i], std::min(out_size[i], out_cap[i]), cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    return DBG_OK;
}

static bool items_in_order(uint64_t n, const uint64_t *in_off, const uint64_t *in_size, const uint64_t *out_off, const uint64_t *out_cap)
{
    for (uint64_t i = 1; i < n; i++)
        if (in_off[i] < in_off[i - 1] + in_size[i - 1] || out_off[i] < out_off[i - 1] + out_cap[i - 1]) return false;
    return true;
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
    ctx->err[0] = 0;
    const bool mono = items_in_order(n, in_off, in_size, out_off, out_cap);
    if (n <= 4 || !mono) return packed_small(ctx, kind, n, h_in, in_off, in_size, h_out, out_off, out_cap, out_size, status);
    uint64_t tot_in = 0;
    for (uint64_t i = 0; i < n; i++) tot_in += in_size[i];
    // ---- waves: contiguous item ranges, each on its own stream and scratch slot (H2D -> kernels -> D2H), so that
    // the copy engines and the SMs overlap
    int nw;
    bool intra = false;
    if (kind == 2) {
        // PNG: a wave should still fill the GPU (>= 64 MB of files)
        nw = (int)std::min<uint64_t>(std::min<uint64_t>(ctx->png_waves, n), std::max<uint64_t>(1, tot_in / (64ull << 20)));
    } else {
        nw = (int)std::min<uint64_t>(ctx->waves, std::max<uint64_t>(1, n / 256));
        if (n < ctx->small_batch) nw = 1;  // small batches: one wave, the intra-stream paths may take its long streams
        if (nw > 1 && ctx->bsplit) {
            // a batch with streams long enough for the block-split path (same rule as bs_classify_kernel) runs
            // as one wave: that path synchronises the host twice, which would serialise concurrent waves
            const uint64_t resident = (uint64_t)ctx->sm_count * ctx->inflate_ctas_per_sm * dbg::INFLATE_WARPS_PER_CTA;
            // (a pipe's batches in flight share the GPU: a member is long relative to all of them)
            const uint64_t thr = std::max<uint64_t>(ctx->bsplit_min_bytes, tot_in * ctx->in_flight_share / resident * ctx->bsplit_factor_q / 4);
            for (uint64_t i = 0; i < n && nw > 1; i++)
                if (in_size[i] >= thr) nw = 1;
        }
        intra = nw == 1;
    }
    std::vector<uint64_t> cuts(nw + 1);
    if (kind == 2) {  // PNG waves of equal bytes (images of one batch may differ a lot in size)
        cuts[0] = 0;
        uint64_t acc = 0, i = 0;
        for (int k = 1; k < nw; k++) {
            const uint64_t want = tot_in * (uint64_t)k / nw;
            while (i < n && acc < want) acc += in_size[i++];
            cuts[k] = std::max<uint64_t>(i, cuts[k - 1] + 1);
            if (cuts[k] > n - (uint64_t)(nw - k)) cuts[k] = n - (uint64_t)(nw - k);
        }
        cuts[nw] = n;
    } else {
        for (int k = 0; k <= nw; k++) cuts[k] = n * (uint64_t)k / nw;
    }
    return packed_waves(ctx, kind, n, h_in, in_off, in_size, h_out, out_off, out_cap, out_size, status, cuts, intra);
}

// One gzip member into memory obtained from the caller's allocator once its size is known (the drop-in decode_gz()
// is built on this: the reference allocates 35 x the input up front, decode_gz.c:245-247). `cap` bounds the output.
extern "C" int dbg_decode_gz_alloc(dbg_ctx *ctx, const uint8_t *in, uint64_t in_size, uint64_t cap, void *(*alloc)(size_t),
                                   uint8_t **out, uint64_t *out_size, uint32_t *good)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (!in || !alloc || !out || !out_size || !good) return DBG_ERR_ARG;
    *out = nullptr;
    *out_size = 0;
    *good = 0;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    Slot &sl = ctx->slot[0];
    CU(ctx->h_desc.reserve(64));
    CU(ctx->d_desc.reserve(64));
    CU(ctx->d_in.reserve(in_size + 128));
    CU(ctx->d_out.reserve(cap + 64));
    uint64_t *hd = (uint64_t *)ctx->h_desc.p, *dd = (uint64_t *)ctx->d_desc.p;
    hd[0] = 0; hd[1] = in_size; hd[2] = 0; hd[3] = cap; hd[4] = 0; hd[5] = 0;
    {
        SlotGuard guard(sl, s);
        CU(cudaMemcpyAsync(dd, hd, 48, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(ctx->d_in.p, in, in_size, cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync((uint8_t *)ctx->d_in.p + in_size, 0, 64, s));
        int rc = inflate_device_slot(ctx, sl, true, true, 1, (const uint8_t *)ctx->d_in.p, dd, dd + 1, (uint8_t *)ctx->d_out.p, dd + 2,
                                     dd + 3, dd + 4, (uint32_t *)(dd + 5), nullptr, s);
        if (rc) return rc;
        CU(cudaMemcpyAsync(hd + 4, dd + 4, 16, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    if (*(uint32_t *)(hd + 5) != 0) return DBG_OK;  // good = 0
    const uint64_t sz = hd[4];
    uint8_t *buf = (uint8_t *)alloc((size_t)(sz ? sz : 1));
    if (!buf) return DBG_ERR_NOMEM;
    if (sz) {
        CU(cudaMemcpyAsync(buf, ctx->d_out.p, sz, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    *out = buf;
    *out_size = sz;
    *good = 1;
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
        if (in_size[i] && !in[i]) {
            set_err(ctx, "batch call: item %llu has a NULL input", (unsigned long long)i);
            return DBG_ERR_ARG;
        }
        in_off[i] = ti;
        ti += align_up(in_size[i] + 16, 16);
        out_off[i] = to;
        to += align_up(out_cap[i], 16);
    }
    if (n <= 4) {
        // a call or two (the scalar drop-in API): no staging copy on the host -- the buffers are used where they are
        // (pageable memory: the driver stages them), item by item
        int rc = DBG_OK;
        for (uint64_t i = 0; i < n && rc == DBG_OK; i++) {
            const uint64_t zero = 0;
            uint64_t osz = 0;
            if (!out[i] && out_cap[i]) {
                set_err(ctx, "batch call: item %llu has a NULL output", (unsigned long long)i);
                return DBG_ERR_ARG;
            }
            rc = packed_small(ctx, kind, 1, in[i], &zero, &in_size[i], out[i], &zero, &out_cap[i], &osz, &status[i]);
            if (out_size) out_size[i] = osz;
        }
        return rc;
    }
    CU(ctx->h_in.reserve(ti + 64));
    CU(ctx->h_out.reserve(to + 64));
    uint8_t *hi = (uint8_t *)ctx->h_in.p;
    for (uint64_t i = 0; i < n; i++) {
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

// -------------------------------------------------------------- multi-device --
// One batch over several GPUs of one box: the items are independent, so the batch is cut into runs
// of consecutive items, the runs are dealt to the devices longest-processing-time first by estimated decode time,
// and every device decodes its runs as waves of its own packed path, driven by one host thread per device. No
// exchange step, no collective. The reference's only provision for concurrency is the thread_id slot
// (inflate.c:22-23, decode_png.c:559-560).
struct dbg_multi {
    std::vector<dbg_ctx *> ctx;
    char err[512] = "";
};

extern "C" void dbg_multi_destroy(dbg_multi *m)
{
    if (!m) return;
    for (dbg_ctx *c : m->ctx) dbg_destroy(c);
    delete m;
}

extern "C" dbg_multi *dbg_multi_create(int n_devices, const int *device_ids)
{
    const int have = dbg_device_count();
    if (have <= 0) {
        if (have == 0) set_err(nullptr, "no CUDA device visible -- this library has no CPU path");
        return nullptr;
    }
    if (n_devices <= 0) n_devices = have;
    dbg_multi *m = new dbg_multi();
    for (int k = 0; k < n_devices; k++) {
        dbg_ctx *c = dbg_create(device_ids ? device_ids[k] : k);
        if (!c) {
            dbg_multi_destroy(m);
            return nullptr;
        }
        m->ctx.push_back(c);
    }
    return m;
}

extern "C" int dbg_multi_device_count(const dbg_multi *m) { return m ? (int)m->ctx.size() : 0; }
extern "C" dbg_ctx *dbg_multi_ctx(dbg_multi *m, int k) { return (m && k >= 0 && k < (int)m->ctx.size()) ? m->ctx[k] : nullptr; }
extern "C" const char *dbg_multi_last_error(const dbg_multi *m) { return m ? m->err : g_err; }

// ---------------------------------------------------------------- sprite sheets --
static int sprite_grid(dbg_ctx *ctx, uint64_t n, uint32_t w, uint32_t h, uint32_t columns, uint64_t sheet_cap, uint32_t *rows_out,
                       uint32_t *cols_out)
{
    if (n == 0 || w == 0 || h == 0 || n > 0x7fffffffull) {
        set_err(ctx, "sprite sheet: no images, or a zero tile size");
        return DBG_ERR_ARG;
    }
    uint64_t cols = columns;
    if (cols == 0)
        for (cols = 1; cols * cols < n; cols++) {}
    if (cols > n) cols = n;
    const uint64_t rows = (n + cols - 1) / cols;
    const uint64_t sheet_w = cols * w, sheet_h = rows * h;
    if (sheet_w > 0xffffffffull || sheet_h > 0xffffffffull || sheet_w * sheet_h > sheet_cap / 4) {
        set_err(ctx, "sprite sheet: %llu x %llu pixels do not fit %llu bytes", (unsigned long long)sheet_w, (unsigned long long)sheet_h,
                (unsigned long long)sheet_cap);
        return DBG_ERR_ARG;
    }
    *rows_out = (uint32_t)rows;
    *cols_out = (uint32_t)cols;
    return DBG_OK;
}

extern "C" int dbg_tile_sprites_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_rgba, const uint64_t *d_rgba_off, uint32_t w, uint32_t h,
                                       uint32_t columns, int offsets_16_aligned, uint8_t *d_sheet, uint64_t sheet_cap, uint32_t *out_rows,
                                       uint32_t *out_columns, void *stream)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (!d_rgba || !d_rgba_off || !d_sheet) {
        set_err(ctx, "dbg_tile_sprites_device: bad arguments");
        return DBG_ERR_ARG;
    }
    uint32_t rows = 0, cols = 0;
    const int rc = sprite_grid(ctx, n, w, h, columns, sheet_cap, &rows, &cols);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const dbg::SpriteBatch b{d_rgba, d_rgba_off, d_sheet, (uint32_t)n, w, h, cols, rows};
    const bool a16 = offsets_16_aligned && (((uintptr_t)d_rgba | (uintptr_t)d_sheet) & 15) == 0;
    CU((cudaError_t)dbg::sprite_launch(b, a16, ctx->sm_count, stream ? (cudaStream_t)stream : ctx->stream));
    ctx->launches += 1;
    if (out_rows) *out_rows = rows;
    if (out_columns) *out_columns = cols;
    return DBG_OK;
}

extern "C" int dbg_tile_sprites(dbg_ctx *ctx, uint64_t n, const uint8_t *const *rgba, uint32_t w, uint32_t h, uint32_t columns,
                                uint8_t *sheet, uint64_t sheet_cap, uint32_t *out_rows, uint32_t *out_columns)
{
    if (!ctx) return DBG_ERR_NO_DEVICE;
    if (!rgba || !sheet) {
        set_err(ctx, "dbg_tile_sprites: bad arguments");
        return DBG_ERR_ARG;
    }
    uint32_t rows = 0, cols = 0;
    int rc = sprite_grid(ctx, n, w, h, columns, sheet_cap, &rows, &cols);
    if (rc) return rc;
    for (uint64_t i = 0; i < n; i++)
        if (!rgba[i]) {
            set_err(ctx, "dbg_tile_sprites: image %llu is NULL", (unsigned long long)i);
            return DBG_ERR_ARG;
        }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t tile = (uint64_t)w * h * 4, tile_al = (tile + 15) & ~15ull, sheet_bytes = (uint64_t)cols * w * rows * h * 4;
    CU(ctx->d_in.reserve(n * tile_al + 64));
    CU(ctx->d_out.reserve(sheet_bytes + 64));
    CU(ctx->h_desc.reserve(n * 8));
    CU(ctx->d_desc.reserve(n * 8));
    uint64_t *h_off = (uint64_t *)ctx->h_desc.p;
    for (uint64_t i = 0; i < n; i++) {
        h_off[i] = i * tile_al;
        CU(cudaMemcpyAsync((uint8_t *)ctx->d_in.p + h_off[i], rgba[i], tile, cudaMemcpyHostToDevice, s));
    }
    CU(cudaMemcpyAsync(ctx->d_desc.p, h_off, n * 8, cudaMemcpyHostToDevice, s));
    rc = dbg_tile_sprites_device(ctx, n, (const uint8_t *)ctx->d_in.p, (const uint64_t *)ctx->d_desc.p, w, h, cols, 1, (uint8_t *)ctx->d_out.p,
                                 sheet_bytes, out_rows, out_columns, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(sheet, ctx->d_out.p, sheet_bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return DBG_OK;
}

// ------------------------------------------------------------------- pipe -----
// Several packed batches in flight on one GPU (see the header): `depth` workers, each a host thread with a context of
// its own, take the submitted batches in order. A worker is inside dbg_decode_batch_packed() for the whole batch, so
// the next batch's uploads and first kernels run under this batch's downloads.
struct dbg_pipe {
    struct Job {
        int64_t ticket;
        int kind;
        uint64_t n;
        const uint8_t *h_in;
        const uint64_t *in_off, *in_size;
        uint8_t *h_out;
        const uint64_t *out_off, *out_cap;
        uint64_t *out_size;
        uint32_t *status;
    };
    std::vector<dbg_ctx *> ctx;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<Job> queue;
    std::vector<std::pair<int64_t, int>> done;  // finished tickets not yet waited for
    std::vector<int64_t> open_tickets;           // submitted, not yet waited for
    int64_t next_ticket = 0;
    int64_t upload_turn = 0;  // the ticket whose uploads may be enqueued: batches upload in ticket order, one after the other
    int in_flight = 0;        // submitted, not yet finished
    bool stop = false;
    struct Turn {
        dbg_pipe *p;
        int64_t ticket;
        bool left;
    };
    static void turn_enter(void *a)
    {
        Turn *t = (Turn *)a;
        std::unique_lock<std::mutex> lk(t->p->mu);
        t->p->cv_done.wait(lk, [&] { return t->p->upload_turn >= t->ticket; });
    }
    static void turn_leave(void *a)
    {
        Turn *t = (Turn *)a;
        if (t->left) return;
        t->left = true;
        {
            std::lock_guard<std::mutex> lk(t->p->mu);
            if (t->p->upload_turn <= t->ticket) t->p->upload_turn = t->ticket + 1;
        }
        t->p->cv_done.notify_all();
    }

    void run(dbg_ctx *c)
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || !queue.empty(); });
                if (queue.empty()) return;
                j = queue.front();
                queue.pop_front();
            }
            Turn turn{this, j.ticket, false};
            c->gate_enter = turn_enter;
            c->gate_leave = turn_leave;
            c->gate_arg = &turn;
            const int rc = dbg_decode_batch_packed(c, j.kind, j.n, j.h_in, j.in_off, j.in_size, j.h_out, j.out_off, j.out_cap,
                                                   j.out_size, j.status);
            c->gate_enter = c->gate_leave = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);  // a batch that took a path without uploads of its own passes the turn on, in order
                cv_done.wait(lk, [&] { return turn.left || upload_turn >= j.ticket; });
            }
            turn_leave(&turn);
            {
                std::lock_guard<std::mutex> lk(mu);
                done.emplace_back(j.ticket, rc);
                in_flight--;
            }
            cv_done.notify_all();
        }
    }
};

extern "C" void dbg_pipe_destroy(dbg_pipe *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;  // workers drain the queue first
    }
    p->cv_job.notify_all();
    for (std::thread &t : p->workers) t.join();
    for (dbg_ctx *c : p->ctx) dbg_destroy(c);
    delete p;
}

extern "C" dbg_pipe *dbg_pipe_create(int device, int depth)
{
    if (depth < 1 || depth > 4) {
        set_err(nullptr, "dbg_pipe_create: depth must be 1..4");
        return nullptr;
    }
    const int have = dbg_device_count();
    if (have <= 0 || device < 0 || device >= have || cudaSetDevice(device) != cudaSuccess) {
        set_err(nullptr, "dbg_pipe_create: no usable device %d -- this library has no CPU path", device);
        return nullptr;
    }
    // Streams are mapped to the device's hardware queues (at most 32) in creation order, and work on streams that share a
    // queue is serialised: two ordinary contexts bring 68 streams, and the waves of one then wait behind the other's
    // kernels (measured: cfg2 through two such contexts 136 ms per step, one blocking call 113 ms). So the wave and upload
    // streams of all the pipe's contexts are made first, back to back, 16 wave streams in all.
    const int w = dbg_ctx::MAX_WAVES / depth;
    std::vector<cudaStream_t> pre((size_t)depth * (w + 1), nullptr);
    for (cudaStream_t &st : pre)
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
            set_err(nullptr, "dbg_pipe_create: %s", cudaGetErrorString(cudaGetLastError()));
            for (cudaStream_t q : pre)
                if (q) cudaStreamDestroy(q);
            return nullptr;
        }
    dbg_pipe *p = new dbg_pipe();
    for (int k = 0; k < depth; k++) {
        dbg_ctx *c = ctx_create(device, pre.data() + (size_t)k * (w + 1), w, pre[(size_t)k * (w + 1) + w]);
        if (!c) {
            for (size_t q = (size_t)(k + 1) * (w + 1); q < pre.size(); q++) cudaStreamDestroy(pre[q]);  // never adopted
            dbg_pipe_destroy(p);
            return nullptr;
        }
        c->in_flight_share = (uint32_t)depth;
        c->ordered_uploads = getenv("DBG_PIPE_ORDERED") != nullptr;  // measured on cfg2: 97.8 ms per step against 95.4 without
        c->waves = std::min(c->waves, w);
        c->png_waves = std::min(c->png_waves, w);
        p->ctx.push_back(c);
    }
    for (int k = 0; k < depth; k++) p->workers.emplace_back([p, k] { p->run(p->ctx[k]); });
    return p;
}

extern "C" int dbg_pipe_depth(const dbg_pipe *p) { return p ? (int)p->ctx.size() : 0; }
extern "C" dbg_ctx *dbg_pipe_ctx(dbg_pipe *p, int k) { return (p && k >= 0 && k < (int)p->ctx.size()) ? p->ctx[k] : nullptr; }

extern "C" int64_t dbg_pipe_submit(dbg_pipe *p, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                                   const uint64_t *in_size, uint8_t *h_out, const uint64_t *out_off, const uint64_t *out_cap,
                                   uint64_t *out_size, uint32_t *status)
{
    if (!p) return DBG_ERR_NO_DEVICE;
    int64_t ticket;
    {
        std::unique_lock<std::mutex> lk(p->mu);
        p->cv_done.wait(lk, [&] { return p->in_flight < (int)p->ctx.size(); });
        ticket = p->next_ticket++;
        p->in_flight++;
        p->open_tickets.push_back(ticket);
        p->queue.push_back(dbg_pipe::Job{ticket, kind, n, h_in, in_off, in_size, h_out, out_off, out_cap, out_size, status});
    }
    p->cv_job.notify_one();
    return ticket;
}

extern "C" int dbg_pipe_wait(dbg_pipe *p, int64_t ticket)
{
    if (!p) return DBG_ERR_NO_DEVICE;
    std::unique_lock<std::mutex> lk(p->mu);
    auto it = std::find(p->open_tickets.begin(), p->open_tickets.end(), ticket);
    if (it == p->open_tickets.end()) return DBG_ERR_ARG;  // never handed out, or waited for already
    p->open_tickets.erase(it);
    for (;;) {
        for (size_t k = 0; k < p->done.size(); k++)
            if (p->done[k].first == ticket) {
                const int rc = p->done[k].second;
                p->done.erase(p->done.begin() + (long)k);
                return rc;
            }
        p->cv_done.wait(lk);
    }
}

// Estimated decode time of an item, in arbitrary units: compressed bytes (a stream that opens with a stored block is a
// plain copy) plus a share of the output (match copying, un-filtering).
static inline uint64_t item_cost(int kind, const uint8_t *p, uint64_t size, uint64_t cap)
{
    return sched_weight(kind, p, size) + cap / (kind == 2 ? 4 : 32) + 4096;
}

// The partition alone (exported so that callers and tests can see it): device_of_item[i] = index of the device that
// gets item i. Returns the number of runs. Runs never split an item; when the items do not lie in index order in the
// arenas every item is a run of its own.
extern "C" int dbg_multi_partition(int n_devices, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                                   const uint64_t *in_size, const uint64_t *out_off, const uint64_t *out_cap,
                                   uint32_t *device_of_item, uint64_t *device_cost)
{
    if (n_devices <= 0 || !in_off || !in_size || !out_off || !out_cap || !device_of_item) return DBG_ERR_ARG;
    std::vector<uint64_t> cost(n);
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; i++) {
        cost[i] = item_cost(kind, h_in ? h_in + in_off[i] : nullptr, h_in ? in_size[i] : 0, out_cap[i]) + (h_in ? 0 : in_size[i]);
        total += cost[i];
    }
    const bool mono = items_in_order(n, in_off, in_size, out_off, out_cap);
    const uint64_t target = std::max<uint64_t>(1, total / ((uint64_t)n_devices * 8));  // ~8 runs per device
    struct Run { uint64_t b, e, cost; };
    std::vector<Run> runs;
    for (uint64_t i = 0; i < n;) {
        Run r{i, i, 0};
        do {
            r.cost += cost[r.e++];
        } while (mono && r.e < n && r.cost + cost[r.e] / 2 < target);
        runs.push_back(r);
        i = r.e;
    }
    std::vector<size_t> by(runs.size());
    std::iota(by.begin(), by.end(), (size_t)0);
    std::stable_sort(by.begin(), by.end(), [&](size_t a, size_t b) { return runs[a].cost > runs[b].cost; });
    std::vector<uint64_t> load(n_devices, 0);
    for (size_t k : by) {
        const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[d] += runs[k].cost;
        for (uint64_t i = runs[k].b; i < runs[k].e; i++) device_of_item[i] = (uint32_t)d;
    }
    if (device_cost)
        for (int d = 0; d < n_devices; d++) device_cost[d] = load[d];
    return (int)runs.size();
}

extern "C" int dbg_decode_batch_packed_multi(dbg_multi *m, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                                             const uint64_t *in_size, uint8_t *h_out, const uint64_t *out_off,
                                             const uint64_t *out_cap, uint64_t *out_size, uint32_t *status, uint32_t *device_of_item)
{
    if (!m || m->ctx.empty()) return DBG_ERR_NO_DEVICE;
    if (n == 0) return DBG_OK;
    if (!h_in || !in_off || !in_size || !h_out || !out_off || !out_cap || !out_size || !status || kind < 0 || kind > 2 ||
        n > 0x7fffffffull) {
        snprintf(m->err, sizeof(m->err), "dbg_decode_batch_packed_multi: bad arguments");
        return DBG_ERR_ARG;
    }
    const int nd = (int)m->ctx.size();
    if (nd == 1) {
        if (device_of_item) memset(device_of_item, 0, n * 4);
        int rc = dbg_decode_batch_packed(m->ctx[0], kind, n, h_in, in_off, in_size, h_out, out_off, out_cap, out_size, status);
        if (rc) memcpy(m->err, m->ctx[0]->err, sizeof(m->err));
        return rc;
    }
    std::vector<uint32_t> dev(n);
    dbg_multi_partition(nd, kind, n, h_in, in_off, in_size, out_off, out_cap, dev.data(), nullptr);
    if (device_of_item) memcpy(device_of_item, dev.data(), n * 4);
    const bool mono = items_in_order(n, in_off, in_size, out_off, out_cap);
    std::vector<int> rcs(nd, DBG_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; d++) {
        th.emplace_back([&, d]() {
            // this device's items, in index order, as a batch of its own; a wave per run of consecutive items
            std::vector<uint64_t> idx, cuts;
            for (uint64_t i = 0; i < n; i++)
                if (dev[i] == (uint32_t)d) {
                    if (idx.empty() || idx.back() + 1 != i) cuts.push_back(idx.size());
                    idx.push_back(i);
                }
            if (idx.empty()) return;
            cuts.push_back(idx.size());
            const uint64_t k = idx.size();
            std::vector<uint64_t> io(k), is(k), oo(k), oc(k), os(k);
            std::vector<uint32_t> st(k);
            for (uint64_t j = 0; j < k; j++) {
                io[j] = in_off[idx[j]];
                is[j] = in_size[idx[j]];
                oo[j] = out_off[idx[j]];
                oc[j] = out_cap[idx[j]];
            }
            dbg_ctx *ctx = m->ctx[d];
            int rc;
            if (cudaSetDevice(ctx->device) != cudaSuccess) {
                rc = DBG_ERR_CUDA;
            } else if (!mono || k <= 4) {
                rc = packed_small(ctx, kind, k, h_in, io.data(), is.data(), h_out, oo.data(), oc.data(), os.data(), st.data());
            } else {
                // a wave per run, at most MAX_WAVES waves per call (a wave never spans bytes of another device's items:
                // its download would overwrite them on the host)
                rc = DBG_OK;
                const size_t nruns = cuts.size() - 1;
                for (size_t r0 = 0; r0 < nruns && rc == DBG_OK; r0 += dbg_ctx::MAX_WAVES) {
                    const size_t r1 = std::min(nruns, r0 + (size_t)dbg_ctx::MAX_WAVES);
                    const uint64_t j0 = cuts[r0], j1 = cuts[r1];
                    std::vector<uint64_t> sub(cuts.begin() + r0, cuts.begin() + r1 + 1);
                    for (auto &c : sub) c -= j0;
                    ctx->err[0] = 0;
                    rc = packed_waves(ctx, kind, j1 - j0, h_in, io.data() + j0, is.data() + j0, h_out, oo.data() + j0, oc.data() + j0,
                                      os.data() + j0, st.data() + j0, sub, kind == 2 || nruns == 1);
                }
            }
            rcs[d] = rc;
            if (rc == DBG_OK)
                for (uint64_t j = 0; j < k; j++) {
                    out_size[idx[j]] = os[j];
                    status[idx[j]] = st[j];
                }
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (rcs[d]) {
            snprintf(m->err, sizeof(m->err), "device %d: %s", m->ctx[d]->device, m->ctx[d]->err);
            return rcs[d];
        }
    return DBG_OK;
}

// ----------------------------------------------------------------------- BMP --
// decode_bmp.c:105-372: header checks + BGRA <-> RGBA swizzle (bmp_kernels.cuh).
static int bmp_launch(dbg_ctx *ctx, bool encode, dbg::BmpBatch b, cudaStream_t s)
{
    const uint64_t n = b.n;
    CU(ctx->slot[0].meta.reserve(n * sizeof(dbg::BmpItem) + (n + 1) * sizeof(uint32_t) + 64));
    b.items = (dbg::BmpItem *)ctx->slot[0].meta.p;
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
