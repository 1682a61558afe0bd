/*
 * decode_bmp.h -- drop-in for the reference's decode_bmp.h (decode_bmp.h:14-36):
 * same three entry points, same argument meaning, served by the batched GPU path
 * (dbg_decode_bmp_batch / dbg_encode_bmp_batch, a batch of one on CUDA device
 * $DBG_DEVICE, default 0). There is no CPU path: without a device decode_BMP
 * reports good = 0 and encode_BMP reports a recipient_size of 0.
 *
 * Behaviour follows decode_bmp.c:
 *   get_BMP_width_height  :53-103  'BM' check, |height|, good = width > 0 && height > 0
 *   decode_BMP            :105-295 header checks, BGRA -> RGBA, bottom-up files flipped
 *   encode_BMP            :297-372 14 + 40 header bytes (top-down), RGBA -> BGRA,
 *                                  *recipient_size = 54 + rgba_size + 1
 * Inputs on which the reference reads or writes out of bounds (pixel data past the
 * end of the file, w*h*4 > out_rgba_values_size, recipient_capacity too small) are
 * rejected here instead.
 */
#ifndef DECODE_BMP_H
#define DECODE_BMP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void get_BMP_width_height(const uint8_t *raw_input, const uint64_t raw_input_size, uint32_t *out_width,
                          uint32_t *out_height, uint8_t *out_good);

void decode_BMP(const uint8_t *raw_input, const uint64_t raw_input_size, uint8_t *out_rgba_values,
                const int64_t out_rgba_values_size, uint8_t *out_good);

void encode_BMP(const uint8_t *rgba, const uint64_t rgba_size, const uint32_t width, const uint32_t height,
                char *recipient, uint32_t *recipient_size, const int64_t recipient_capacity);

#ifdef __cplusplus
}
#endif

#endif /* DECODE_BMP_H */
