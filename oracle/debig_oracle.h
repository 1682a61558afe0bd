/* oracle/debig_oracle.h -- CPU restatement of the reference decode path.
 * TEST INFRASTRUCTURE ONLY; see debig_oracle.c. */
#ifndef DEBIG_ORACLE_H
#define DEBIG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* in_avail >= in_size: bytes that may be read past the declared size (the
 * reference's bit reader over-reads, inflate.c:252-256). */
void oracle_inflate(const uint8_t *in, uint64_t in_size, uint64_t in_avail, uint8_t *out, uint64_t cap,
                    uint64_t *out_size, uint32_t *good);
void oracle_stats(uint64_t *out8);
void oracle_decode_gz(const uint8_t *in, uint64_t in_size, uint8_t *out, uint64_t cap, uint64_t *out_size,
                      uint32_t *good);
void oracle_png_dims(const uint8_t *in, uint64_t in_size, uint32_t *w, uint32_t *h, uint8_t *good);
void oracle_decode_png(const uint8_t *file, uint64_t size, uint8_t *out, uint64_t rgba_size, int rgb_as_reference,
                       uint8_t *good);
void oracle_bmp_dims(const uint8_t *in, uint64_t size, uint32_t *w, uint32_t *h, uint8_t *good);
void oracle_decode_bmp(const uint8_t *in, uint64_t size, uint8_t *out, int64_t out_size, uint8_t *good);
void oracle_encode_bmp(const uint8_t *rgba, uint64_t rgba_size, uint32_t w, uint32_t h, uint8_t *out,
                       uint32_t *out_size, int64_t cap);
#ifdef __cplusplus
}
#endif
#endif
