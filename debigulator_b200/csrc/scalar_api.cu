// scalar_api.cu -- the reference's header-level API (include/inflate.h,
// decode_png.h, decode_gz.h) on top of the batched GPU path: every call is a
// batch of one through dbg_*_batch(). Mirrors the reference's call contract
// (argument checks, out_good convention, per-slot init) so existing callers
// of inflate.c / decode_png.c / decode_gz.c can relink unchanged.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/debigulator_b200.h"
#include "../../include/decode_bmp.h"
#include "../../include/decode_gz.h"
#include "../../include/decode_png.h"
#include "../../include/inflate.h"

#define DBG_MAX_SLOTS 10  // INFLATE_MAX_THREADS inflate.c:22, PNG_DECODER_MAX_THREADS decode_png.c:559

static dbg_ctx *g_ctx[DBG_MAX_SLOTS];
static uint32_t g_png_init[DBG_MAX_SLOTS];
static uint32_t g_png_wm[DBG_MAX_SLOTS];
static void *(*g_gz_malloc)(size_t) = nullptr;

static dbg_ctx *slot_ctx(uint32_t slot)
{
    if (slot >= DBG_MAX_SLOTS) return nullptr;
    if (!g_ctx[slot]) {
        const char *dev = getenv("DBG_DEVICE");
        g_ctx[slot] = dbg_create(dev ? atoi(dev) : 0);
    }
    return g_ctx[slot];
}

extern "C" void inflate_init(void *(*)(uint64_t), void *(*)(void *, int, uint64_t), void *(*)(void *, const void *, uint64_t),
                             const uint32_t thread_id)
{
    (void)slot_ctx(thread_id);  // inflate.c:40-56: allocate the slot's state
}

extern "C" void inflate_destroy(void (*)(void *), const uint32_t thread_id)
{
    if (thread_id < DBG_MAX_SLOTS && g_ctx[thread_id]) {
        dbg_destroy(g_ctx[thread_id]);
        g_ctx[thread_id] = nullptr;
    }
}

extern "C" void inflate(uint8_t const *recipient, const uint64_t recipient_size, uint64_t *final_recipient_size, uint8_t *,
                        const uint64_t, uint8_t const *compressed_input, const uint64_t compressed_input_size,
                        uint32_t *out_good, const uint32_t thread_id)
{
    if (!out_good) return;
    *out_good = 0;
    // inflate.c:797-824
    if (recipient == nullptr || final_recipient_size == nullptr || compressed_input == nullptr) return;
    dbg_ctx *ctx = slot_ctx(thread_id);
    if (!ctx) return;
    const uint8_t *in[1] = {compressed_input};
    uint8_t *out[1] = {(uint8_t *)recipient};
    uint64_t in_size[1] = {compressed_input_size}, cap[1] = {recipient_size}, sz[1] = {0};
    uint32_t good[1] = {0};
    if (dbg_inflate_batch(ctx, 1, in, in_size, out, cap, sz, good) != DBG_OK) return;
    if (good[0]) *final_recipient_size = sz[0];
    else if (recipient_size >= compressed_input_size && compressed_input_size >= 5) *final_recipient_size = 0;  // :852
    *out_good = good[0];
}

extern "C" void decode_png_init(void *(*)(uint64_t), void (*)(void *), void *(*)(void *, int, uint64_t),
                                void *(*)(void *, const void *, uint64_t), const uint32_t dpng_working_memory_size,
                                const uint32_t thread_id)
{
    if (thread_id >= DBG_MAX_SLOTS) return;
    if (g_png_init[thread_id]) return;  // decode_png.c:580-593
    g_png_init[thread_id] = 1;
    g_png_wm[thread_id] = dpng_working_memory_size;
    (void)slot_ctx(thread_id);
}

extern "C" void decode_png_deinit(const uint32_t thread_id)
{
    if (thread_id >= DBG_MAX_SLOTS) return;
    g_png_init[thread_id] = 0;
}

static uint32_t rd_be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

extern "C" void decode_png_get_width_height(const uint8_t *compressed_input, const uint64_t compressed_input_size,
                                            uint32_t *out_width, uint32_t *out_height, uint8_t *out_good)
{
    *out_width = 0;
    *out_height = 0;
    *out_good = 0;
    if (compressed_input_size < 28) return;                                      // decode_png.c:627
    if (memcmp(compressed_input + 1, "PNG", 3) != 0) return;                     // :644-655
    *out_width = rd_be32(compressed_input + 16);                                 // :662-671
    *out_height = rd_be32(compressed_input + 20);
    *out_good = 1;
}

extern "C" void decode_png(const uint8_t *compressed_input, const uint64_t compressed_input_size,
                           const uint8_t *out_rgba_values, const uint64_t rgba_values_size, const uint32_t thread_id,
                           uint8_t *out_good)
{
    if (!out_good) return;
    *out_good = 0;
    if (thread_id >= DBG_MAX_SLOTS || !g_png_init[thread_id]) return;            // decode_png.c:691-700
    if (!compressed_input || !out_rgba_values) return;
    dbg_ctx *ctx = slot_ctx(thread_id);
    if (!ctx) return;
    // working-memory limit chosen at init, decode_png.c:1060-1081
    uint32_t w = 0, h = 0;
    uint8_t ok = 0;
    decode_png_get_width_height(compressed_input, compressed_input_size, &w, &h, &ok);
    if (ok) {
        uint64_t need = (uint64_t)w * h * 4 + h + 1 + 3000000ull;
        if (need > g_png_wm[thread_id]) return;
    }
    const uint8_t *in[1] = {compressed_input};
    uint8_t *out[1] = {(uint8_t *)out_rgba_values};
    uint64_t in_size[1] = {compressed_input_size}, cap[1] = {rgba_values_size};
    uint8_t good[1] = {0};
    if (dbg_decode_png_batch(ctx, 1, in, in_size, out, cap, good) != DBG_OK) return;
    *out_good = good[0];
}

extern "C" void init_PNG_decoder(void *(*)(size_t))
{
    decode_png_init(nullptr, nullptr, nullptr, nullptr, 0xffffffffu, 0);
}
extern "C" void get_PNG_width_height(const uint8_t *compressed_input, const uint64_t compressed_input_size,
                                     uint32_t *out_width, uint32_t *out_height, uint32_t *out_good)
{
    uint8_t g = 0;
    decode_png_get_width_height(compressed_input, compressed_input_size, out_width, out_height, &g);
    *out_good = g;
}
extern "C" void decode_PNG(const uint8_t *compressed_input, const uint64_t compressed_input_size,
                           const uint8_t *out_rgba_values, const uint64_t rgba_values_size, uint32_t *out_good)
{
    uint8_t g = 0;
    decode_png(compressed_input, compressed_input_size, out_rgba_values, rgba_values_size, 0, &g);
    *out_good = g;
}

extern "C" void init_decode_gz(void *(*malloc_funcptr)(size_t), void *(*)(void *, int, size_t),
                               void *(*)(void *, const void *, size_t))
{
    g_gz_malloc = malloc_funcptr;
    (void)slot_ctx(0);
}

extern "C" DecodedData *decode_gz(uint8_t *compressed_bytes, uint32_t compressed_bytes_size)
{
    if (!g_gz_malloc) return nullptr;  // decode_gz.c:105-113
    DecodedData *r = (DecodedData *)g_gz_malloc(sizeof(DecodedData));
    if (!r) return nullptr;
    r->data = nullptr;
    r->data_size = 0;
    r->good = 0;
    if (!compressed_bytes || compressed_bytes_size < 18) return r;
    dbg_ctx *ctx = slot_ctx(0);
    if (!ctx) return r;
    // decode_gz.c:245 sizes the output as left*35 + 1,000,000. ISIZE (mod 2^32) from the trailer raises that bound when
    // it is believable: DEFLATE cannot expand by more than 1032:1, so a larger claim is not honoured (an 18-byte file
    // must not make this call reserve 4 GiB). The buffer handed to the caller is allocated once the size is known.
    uint64_t cap = (uint64_t)compressed_bytes_size * 35 + 1000000ull;
    const uint8_t *t = compressed_bytes + compressed_bytes_size - 4;
    const uint64_t isize = (uint64_t)t[0] | ((uint64_t)t[1] << 8) | ((uint64_t)t[2] << 16) | ((uint64_t)t[3] << 24);
    if (isize > cap && isize <= (uint64_t)compressed_bytes_size * 1032 + 1024) cap = isize;
    if (cap > 0xffffffffull - 2048) cap = 0xffffffffull - 2048;  // the decoder's own limit (inflate_core.h)
    uint8_t *buf = nullptr;
    uint64_t sz = 0;
    uint32_t good = 0;
    if (dbg_decode_gz_alloc(ctx, compressed_bytes, compressed_bytes_size, cap, g_gz_malloc, &buf, &sz, &good) != DBG_OK || !good) return r;
    r->data = (char *)buf;
    r->data_size = (uint32_t)sz;
    r->good = 1;
    return r;
}

// ----------------------------------------------------------------------- BMP --
extern "C" void get_BMP_width_height(const uint8_t *raw_input, const uint64_t raw_input_size, uint32_t *out_width,
                                     uint32_t *out_height, uint8_t *out_good)
{
    *out_good = 0;
    if (!raw_input || raw_input_size < 54) return;  // the reference asserts (decode_bmp.c:69-80)
    if (raw_input[0] != 'B' || raw_input[1] != 'M') return;  // :84-99 (width / height left untouched)
    int32_t w, h;
    memcpy(&w, raw_input + 18, 4);
    memcpy(&h, raw_input + 22, 4);
    *out_height = (uint32_t)(h < 0 ? -(int64_t)h : h);  // :101-105
    *out_width = (uint32_t)w;
    *out_good = *out_width > 0 && *out_height > 0;
}

extern "C" void decode_BMP(const uint8_t *raw_input, const uint64_t raw_input_size, uint8_t *out_rgba_values,
                           const int64_t out_rgba_values_size, uint8_t *out_good)
{
    if (!out_good) return;
    *out_good = 0;
    if (!raw_input || !out_rgba_values || out_rgba_values_size < 0) return;
    dbg_ctx *ctx = slot_ctx(0);
    if (!ctx) return;
    const uint8_t *in[1] = {raw_input};
    uint8_t *out[1] = {out_rgba_values};
    uint64_t in_size[1] = {raw_input_size}, cap[1] = {(uint64_t)out_rgba_values_size};
    uint8_t good[1] = {0};
    if (dbg_decode_bmp_batch(ctx, 1, in, in_size, out, cap, nullptr, nullptr, good) != DBG_OK) return;
    *out_good = good[0];
}

extern "C" void encode_BMP(const uint8_t *rgba, const uint64_t rgba_size, const uint32_t width, const uint32_t height,
                           char *recipient, uint32_t *recipient_size, const int64_t recipient_capacity)
{
    if (!recipient_size) return;
    *recipient_size = 0;
    if (!rgba || !recipient || recipient_capacity < 0) return;
    dbg_ctx *ctx = slot_ctx(0);
    if (!ctx) return;
    const uint8_t *in[1] = {rgba};
    uint8_t *out[1] = {(uint8_t *)recipient};
    uint64_t in_size[1] = {rgba_size}, cap[1] = {(uint64_t)recipient_capacity}, sz[1] = {0};
    uint32_t w[1] = {width}, h[1] = {height}, st[1] = {0};
    if (dbg_encode_bmp_batch(ctx, 1, in, in_size, w, h, out, cap, sz, st) != DBG_OK || st[0]) return;
    *recipient_size = (uint32_t)sz[0];
}
