/*
 * decode_png.h -- drop-in for the reference's src/decode_png.h (:43-50
 * decode_png_init, :52-53 decode_png_deinit, :69-75
 * decode_png_get_width_height, :96-103 decode_png), backed by the sm_100a CUDA
 * kernels in libdebigulator_b200.so, plus the older decode_PNG-style names the
 * reference's own example still calls (hellopng.c:154,176,200).
 *
 * Output is always RGBA8, width*height*4 bytes, rows top-down, no padding.
 * Unlike the reference the input buffer is NOT overwritten. Colour type 2
 * (RGB) decodes correctly here (the reference's expansion is broken,
 * decode_png.c:1509-1536). No CPU fallback: without a CUDA device *out_good
 * is 0.
 */
#ifndef DECODE_PNG_H
#define DECODE_PNG_H

#include <stddef.h>
#include <stdint.h>

#include "inflate.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Must run once per thread_id slot (0..9) before decode_png(). The function
 * pointers are kept for source compatibility. dpng_working_memory_size keeps
 * its reference meaning as a limit: an image needing more than
 * width*height*4 + height + 1 + 3,000,000 bytes is rejected
 * (decode_png.c:1060-1081). */
void decode_png_init(
    void *(*malloc_funcptr)(uint64_t size),
    void (*arg_free_funcptr)(void *),
    void *(*arg_memset_funcptr)(void *str, int c, uint64_t n),
    void *(*arg_memcpy_func)(void *dest, const void *src, uint64_t n),
    const uint32_t dpng_working_memory_size,
    const uint32_t thread_id);

void decode_png_deinit(const uint32_t thread_id);

/* Reads width and height from the first 28 bytes (host only, no GPU work). */
void decode_png_get_width_height(
    const uint8_t *compressed_input,
    const uint64_t compressed_input_size,
    uint32_t *out_width,
    uint32_t *out_height,
    uint8_t *out_good);

/* rgba_values_size must equal width*height*4. *out_good: 1 success, 0 failure. */
void decode_png(
    const uint8_t *compressed_input,
    const uint64_t compressed_input_size,
    const uint8_t *out_rgba_values,
    const uint64_t rgba_values_size,
    const uint32_t thread_id,
    uint8_t *out_good);

/* Legacy names (call sites only survive in the reference: hellopng.c:200,
 * :154-164, :176-186). Thin wrappers over slot 0. */
void init_PNG_decoder(void *(*malloc_funcptr)(size_t size));
void get_PNG_width_height(
    const uint8_t *compressed_input,
    const uint64_t compressed_input_size,
    uint32_t *out_width,
    uint32_t *out_height,
    uint32_t *out_good);
void decode_PNG(
    const uint8_t *compressed_input,
    const uint64_t compressed_input_size,
    const uint8_t *out_rgba_values,
    const uint64_t rgba_values_size,
    uint32_t *out_good);

#ifdef __cplusplus
}
#endif

#endif /* DECODE_PNG_H */
