/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin driver around the UNMODIFIED reference sources, which are compiled in
 * place from /root/reference/src by oracle/Makefile into oracle/_ref/libref.so.
 * Nothing from the reference tree is copied into this repository: this file
 * only declares call sequences against the reference's public headers
 * (inflate.h:22-60, decode_png.h:43-103), which are found with -I at build time.
 *
 * What it adds on top of the reference entry points:
 *   ref_init / ref_inflate / ref_decode_png  : one-call wrappers that own the
 *       scratch sizing rules the reference needs: silent build,
 *       3,250,000 B inflate scratch, PNG working memory est + 3,250,000.
 *   ref_decode_gz : decode_gz.c itself does not compile against inflate.h
 *       (decode_gz.c:15,256-272 use an older arity), so its header walk
 *       (decode_gz.c:123-233, silent build => FCOMMENT is NOT skipped) is
 *       restated here and handed to the reference inflate() with
 *       size = remaining - 8 exactly as decode_gz.c:270 does.
 *   ref_stb_png   : synthetic PNG writer = the reference's vendored
 *       stb_write.h (stb_write.h:1128 stbi_write_png_to_mem), included from
 *       the reference tree, used as a test-vector generator only.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "inflate.h"
#include "decode_png.h"

#define STB_IMAGE_WRITE_IMPLEMENTATION
#define STBI_WRITE_NO_STDIO
#define STB_IMAGE_WRITE_STATIC
#include "stb_write.h"

#define REF_INFLATE_SCRATCH 3250000ull

static void *m64(uint64_t n) { return malloc((size_t)n); }
static void *ms64(void *p, int c, uint64_t n) { return memset(p, c, (size_t)n); }
static void *mc64(void *d, const void *s, uint64_t n) { return memmove(d, s, (size_t)n); }

static int g_inflate_ready = 0;
static int g_png_ready = 0;
static uint64_t g_png_wm = 0;
static uint8_t *g_scratch = NULL;

void ref_init(void)
{
    if (!g_inflate_ready && !g_png_ready) {
        /* decode_png_init calls inflate_init for the same slot, so only one of
         * the two may run first; plain inflate use initialises lazily here. */
        inflate_init(m64, ms64, mc64, 0);
        g_inflate_ready = 1;
    }
    if (!g_scratch) g_scratch = (uint8_t *)malloc(REF_INFLATE_SCRATCH);
}

/* raw DEFLATE through the reference inflate() (inflate.c:786).
 * `in` must have >= 4 readable bytes after in_size (bit reader over-read,
 * inflate.c:252-256) and `out` >= 1032 bytes of slack after cap (Q4). */
void ref_inflate(const uint8_t *in, uint64_t in_size, uint8_t *out,
                 uint64_t cap, uint64_t *out_size, uint32_t *good)
{
    ref_init();
    *good = 0;
    *out_size = 0;
    inflate(out, cap, out_size, g_scratch, REF_INFLATE_SCRATCH, in, in_size,
            good, 0);
}

/* gzip member: header walk restated from decode_gz.c:123-233 (silent build). */
void ref_decode_gz(const uint8_t *in, uint32_t in_size, uint8_t *out,
                   uint64_t cap, uint64_t *out_size, uint32_t *good)
{
    *good = 0;
    *out_size = 0;
    if (in == NULL || in_size < 10) return;            /* decode_gz.c:117-129 */
    if (in[0] != 31 || in[1] != 139) return;           /* :138-146 */
    if (in[2] != 8) return;                            /* :148-154 */
    uint32_t at = 10, left = in_size - 10;
    if ((in[3] >> 3) & 1) {                            /* FNAME :195-214 */
        uint32_t n = 0;
        while (in[at + n] != 0 && n < left) n++;       /* :46-60 */
        at += n + 1;
        left -= n + 1;
    }
    /* FCOMMENT is only consumed in verbose builds (:223-233) */
    ref_inflate(in + at, (uint64_t)left - 8, out, cap, out_size, good);
}

/* PNG through the reference decode_png() (decode_png.c:683). The input is
 * copied first because decode_png compacts IDAT payloads into it (:1285-1291). */
void ref_png_dims(const uint8_t *in, uint64_t in_size, uint32_t *w, uint32_t *h,
                  uint8_t *good)
{
    decode_png_get_width_height(in, in_size, w, h, good);
}

void ref_decode_png(const uint8_t *in, uint64_t in_size, uint8_t *out_rgba,
                    uint64_t rgba_size, uint8_t *good)
{
    uint32_t w = 0, h = 0;
    uint8_t ok = 0;
    *good = 0;
    decode_png_get_width_height(in, in_size, &w, &h, &ok);
    if (!ok) return;
    uint64_t need = (uint64_t)w * h * 4 + h + 1 + REF_INFLATE_SCRATCH;
    if (need >= 0xFFFFFFFFull) return;   /* decode_png_init takes a uint32 size */
    if (g_png_ready && need > g_png_wm) {
        decode_png_deinit(0);
        g_png_ready = 0;
    }
    if (!g_png_ready) {
        /* decode_png_init -> inflate_init on the same slot; harmless if the
         * slot is already initialised in a no-assert build (inflate.c:50). */
        decode_png_init(m64, free, ms64, mc64, (uint32_t)need, 0);
        g_png_ready = 1;
        g_inflate_ready = 1;
        g_png_wm = need;
    }
    uint8_t *copy = (uint8_t *)malloc(in_size + 64);
    memcpy(copy, in, in_size);
    memset(copy + in_size, 0, 64);
    decode_png(copy, in_size, out_rgba, rgba_size, 0, good);
    free(copy);
}

/* Synthetic PNG generator (BASELINE configs 3-4): stb_write.h:1128.
 * filter = -1 adaptive, 0..4 forced (stb_write.h:253). Returns malloc'd bytes. */
uint8_t *ref_stb_png(const uint8_t *pixels, int w, int h, int comp, int filter,
                     int *out_len)
{
    stbi_write_force_png_filter = filter;
    return stbi_write_png_to_mem(pixels, w * comp, w, h, comp, out_len);
}

/* zlib stream straight from stb's compressor (stb_write.h:895). */
uint8_t *ref_stb_zlib(uint8_t *data, int len, int quality, int *out_len)
{
    return stbi_zlib_compress(data, len, out_len, quality);
}

void ref_free(void *p) { free(p); }
