/*
 * decode_gz.h -- drop-in for the reference's src/decode_gz.h (:23-27
 * DecodedData, :29-34 init_decode_gz, :36-38 decode_gz), backed by the sm_100a
 * CUDA decoder in libdebigulator_b200.so.
 *
 * Same framing rules as the reference's silent build (decode_gz.c:131-233):
 * magic 1F 8B, CM 8, FNAME skipped, FCOMMENT / FEXTRA / FHCRC not handled,
 * CRC32 and ISIZE not verified. Two fixes a caller can rely on: data_size is
 * set (the reference never sets it), and on failure the struct is fully
 * initialised (good = 0, data = NULL, data_size = 0; Q15 in SURVEY.md).
 * The output buffer comes from the injected malloc; the caller frees `data`
 * and the struct. No CPU fallback: without a CUDA device good is 0.
 */
#ifndef DECODE_GZ_H
#define DECODE_GZ_H

#include <stddef.h>
#include <stdint.h>

#include "inflate.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DecodedData {
    char *data;
    uint32_t data_size;
    uint32_t good;
} DecodedData;

void init_decode_gz(
    void *(*malloc_funcptr)(size_t size),
    void *(*arg_memset_func)(void *str, int c, size_t n),
    void *(*arg_memcpy_func)(void *dest, const void *src, size_t n));

/* Returns NULL if init_decode_gz() has not been called (decode_gz.c:105-113). */
DecodedData *decode_gz(
    uint8_t *compressed_bytes,
    uint32_t compressed_bytes_size);

#ifdef __cplusplus
}
#endif

#endif /* DECODE_GZ_H */
