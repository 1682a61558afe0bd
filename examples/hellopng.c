/*
 * examples/hellopng.c -- a fresh example main in the role of the reference's src/hellopng.c
 * (which no longer builds against its own headers): decode the PNG files
 * named on the command line through the drop-in API of include/decode_png.h and print the
 * summary the reference's README shows (README.md:41-47).
 *
 *   gcc -std=c99 -Iinclude examples/hellopng.c -Ldebigulator_b200 -ldebigulator_b200 \
 *       -Wl,-rpath,$PWD/debigulator_b200 -o hellopng
 *   ./hellopng tests/golden/gimp_test.png
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "decode_png.h"

static void *malloc64(uint64_t n) { return malloc((size_t)n); }
static void *memset64(void *p, int c, uint64_t n) { return memset(p, c, (size_t)n); }
static void *memcpy64(void *d, const void *s, uint64_t n) { return memcpy(d, s, (size_t)n); }

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s file.png [...]\n", argv[0]);
        return 2;
    }
    decode_png_init(malloc64, free, memset64, memcpy64, 0xffffffffu, 0);
    int failures = 0;
    for (int a = 1; a < argc; a++) {
        const char *name = strrchr(argv[a], '/') ? strrchr(argv[a], '/') + 1 : argv[a];
        printf("Inspecting file: %s\n", name);
        FILE *f = fopen(argv[a], "rb");
        if (!f) {
            printf("could not open %s\n", argv[a]);
            failures++;
            continue;
        }
        fseek(f, 0, SEEK_END);
        long size = ftell(f);
        fseek(f, 0, SEEK_SET);
        uint8_t *buf = (uint8_t *)malloc((size_t)size + 16);
        size_t got = fread(buf, 1, (size_t)size, f);
        fclose(f);
        printf("bytes read from raw file: %zu\n", got);
        uint32_t w = 0, h = 0;
        uint8_t good = 0;
        decode_png_get_width_height(buf, got, &w, &h, &good);
        if (!good) {
            printf("finished decode_PNG, result was: FAILURE (not a PNG header)\n");
            failures++;
            free(buf);
            continue;
        }
        uint64_t rgba_size = (uint64_t)w * h * 4;
        uint8_t *rgba = (uint8_t *)malloc((size_t)rgba_size);
        decode_png(buf, got, rgba, rgba_size, 0, &good);
        printf("finished decode_PNG, result was: %s\n", good ? "SUCCESS" : "FAILURE");
        if (good) {
            uint64_t sum[4] = {0, 0, 0, 0};
            for (uint64_t i = 0; i < rgba_size; i++) sum[i & 3] += rgba[i];
            uint64_t px = (uint64_t)w * h;
            printf("rgba values in image: %llu\n", (unsigned long long)rgba_size);
            printf("pixels in image (info from image header): %llu\n", (unsigned long long)px);
            printf("image width: %u\n", w);
            printf("image height: %u\n", h);
            printf("average pixel: [%llu,%llu,%llu,%llu]\n", (unsigned long long)(sum[0] / px), (unsigned long long)(sum[1] / px),
                   (unsigned long long)(sum[2] / px), (unsigned long long)(sum[3] / px));
        } else {
            failures++;
        }
        free(rgba);
        free(buf);
    }
    decode_png_deinit(0);
    return failures ? 1 : 0;
}
