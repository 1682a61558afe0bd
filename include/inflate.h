/*
 * inflate.h -- drop-in for the reference's src/inflate.h (same three entry
 * points, same signatures: inflate.h:22-26 inflate_init, :28-30
 * inflate_destroy, :51-60 inflate), backed by the sm_100a CUDA decoder in
 * libdebigulator_b200.so. A call decodes one raw DEFLATE stream on the GPU
 * (a batch of one; use debigulator_b200.h for real batches). No CPU fallback:
 * without a CUDA device *out_good is 0.
 *
 * Differences a caller can observe: temp_working_memory is never touched (the
 * tables live in shared memory); bytes of `recipient` past
 * *final_recipient_size are left untouched (the reference dirties up to 774 of
 * them, inflate.c:1861-1870); an output overflow fails the call instead of
 * writing past recipient_size.
 */
#ifndef INFLATE_H
#define INFLATE_H

#include <inttypes.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The injected allocator / memset / memcpy are accepted for source
 * compatibility; only thread_id (0..9, one host thread per slot) matters. */
void inflate_init(
    void *(*malloc_funcptr)(uint64_t size),
    void *(*arg_memset_func)(void *str, int c, uint64_t n),
    void *(*arg_memcpy_func)(void *dest, const void *src, uint64_t n),
    const uint32_t thread_id);

void inflate_destroy(
    void (*free_funcptr)(void *to_free),
    const uint32_t thread_id);

/*
 * recipient / recipient_size       output buffer and its capacity (must be
 *                                  >= compressed_input_size, as in the reference)
 * final_recipient_size             receives the decompressed size
 * temp_working_memory(_size)       ignored
 * compressed_input(_size)          raw DEFLATE stream, at least 5 bytes
 * out_good                         set to 1 on success, 0 on failure
 */
void inflate(
    uint8_t const *recipient,
    const uint64_t recipient_size,
    uint64_t *final_recipient_size,
    uint8_t *temp_working_memory,
    const uint64_t temp_working_memory_size,
    uint8_t const *compressed_input,
    const uint64_t compressed_input_size,
    uint32_t *out_good,
    const uint32_t thread_id);

#ifdef __cplusplus
}
#endif

#endif /* INFLATE_H */
