/*
 * examples/hellopipe.c -- streaming batches through the decoder with several packed calls in flight
 * (dbg_pipe_*, include/debigulator_b200.h; the reference's counterpart is a loop over decode_gz(),
 * decode_gz.c:123). Reads one gzip file, makes batches of `members` copies of it in two arenas and
 * pushes `batches` of them through a pipe of depth 2; prints the throughput and checks every size.
 *
 *   gcc -std=c99 -Iinclude examples/hellopipe.c -Ldebigulator_b200 -ldebigulator_b200 \
 *       -Wl,-rpath,$PWD/debigulator_b200 -o hellopipe
 *   ./hellopipe tests/golden/gzipsample.gz 2048 8
 *
 * (The arenas here come from malloc; pin them -- cudaHostAlloc / cudaHostRegister -- for the copies to run at link speed.)
 */
#define _POSIX_C_SOURCE 199309L /* clock_gettime */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "debigulator_b200.h"

typedef struct {
    uint8_t *h_out;
    uint64_t *out_size;
    uint32_t *status;
} Arena;

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s file.gz [members per batch] [batches]\n", argv[0]);
        return 2;
    }
    const uint64_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 2048;
    const int batches = argc > 3 ? atoi(argv[3]) : 8;
    FILE *f = fopen(argv[1], "rb");
    if (!f) {
        printf("could not open %s\n", argv[1]);
        return 1;
    }
    fseek(f, 0, SEEK_END);
    const uint64_t fsize = (uint64_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *file = (uint8_t *)malloc(fsize);
    if (fread(file, 1, fsize, f) != fsize) return 1;
    fclose(f);
    /* ISIZE: the last four bytes of a gzip member */
    const uint64_t isize = (uint64_t)file[fsize - 4] | ((uint64_t)file[fsize - 3] << 8) | ((uint64_t)file[fsize - 2] << 16) |
                           ((uint64_t)file[fsize - 1] << 24);
    const uint64_t in_stride = (fsize + 16 + 15) / 16 * 16, out_stride = (isize + fsize + 64 + 15) / 16 * 16;
    uint8_t *h_in = (uint8_t *)calloc(n * in_stride + 64, 1);
    uint64_t *in_off = (uint64_t *)malloc(n * 8), *in_size = (uint64_t *)malloc(n * 8);
    uint64_t *out_off = (uint64_t *)malloc(n * 8), *out_cap = (uint64_t *)malloc(n * 8);
    for (uint64_t i = 0; i < n; i++) {
        in_off[i] = i * in_stride;
        in_size[i] = fsize;
        out_off[i] = i * out_stride;
        out_cap[i] = out_stride;
        memcpy(h_in + in_off[i], file, fsize);
    }
    Arena a[2];
    for (int k = 0; k < 2; k++) {
        a[k].h_out = (uint8_t *)malloc(n * out_stride + 64);
        a[k].out_size = (uint64_t *)calloc(n, 8);
        a[k].status = (uint32_t *)calloc(n, 4);
    }
    dbg_pipe *p = dbg_pipe_create(0, 2);
    if (!p) {
        printf("dbg_pipe_create failed: %s\n", dbg_last_error(NULL));
        return 1;
    }
    int64_t ticket[2] = {-1, -1};
    int bad = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int b = 0; b < batches + 2; b++) {
        const int k = b & 1;
        if (ticket[k] >= 0) { /* the batch that used this arena two submissions ago */
            if (dbg_pipe_wait(p, ticket[k]) != DBG_OK) bad++;
            for (uint64_t i = 0; i < n; i++)
                if (a[k].status[i] != 0 || a[k].out_size[i] != isize) bad++;
            ticket[k] = -1;
        }
        if (b < batches)
            ticket[k] = dbg_pipe_submit(p, 1 /* gzip */, n, h_in, in_off, in_size, a[k].h_out, out_off, out_cap, a[k].out_size, a[k].status);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    printf("%d batches of %llu members (%llu bytes each) in %.1f ms: %.2f GB/s of output, %d failures\n", batches,
           (unsigned long long)n, (unsigned long long)isize, s * 1e3, (double)batches * (double)n * (double)isize / s / 1e9, bad);
    dbg_pipe_destroy(p);
    return bad != 0;
}
