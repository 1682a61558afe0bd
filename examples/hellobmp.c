/*
 * examples/hellobmp.c -- a fresh example main in the role of the reference's src/hellobmp.c: read a
 * 32-bit .bmp through the drop-in API of include/decode_bmp.h (get_BMP_width_height, decode_BMP),
 * print what was found, encode the pixels back with encode_BMP and check that the round trip is exact.
 *
 *   gcc -std=c99 -Iinclude examples/hellobmp.c -Ldebigulator_b200 -ldebigulator_b200 \
 *       -Wl,-rpath,$PWD/debigulator_b200 -o hellobmp
 *   ./hellobmp tests/golden/fs_psychologist.bmp [out.bmp]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "decode_bmp.h"

int main(int argc, char **argv)
{
    if (argc < 2 || argc > 3) {
        fprintf(stderr, "usage: %s file.bmp [reencoded.bmp]\n", argv[0]);
        return 2;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) {
        printf("could not open %s\n", argv[1]);
        return 1;
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
    get_BMP_width_height(buf, got, &w, &h, &good);
    if (!good) {
        printf("get_BMP_width_height result was: FAILURE\n");
        return 1;
    }
    printf("image width: %u\nimage height: %u\n", w, h);
    const uint64_t rgba_size = (uint64_t)w * h * 4;
    uint8_t *rgba = (uint8_t *)malloc(rgba_size);
    decode_BMP(buf, got, rgba, (int64_t)rgba_size, &good);
    printf("decode_BMP result was: %s\n", good ? "SUCCESS" : "FAILURE");
    if (!good) return 1;
    unsigned long long sum[4] = {0, 0, 0, 0};
    for (uint64_t i = 0; i < rgba_size; i++) sum[i & 3] += rgba[i];
    printf("average pixel: [%llu,%llu,%llu,%llu]\n", sum[0] / ((unsigned long long)w * h), sum[1] / ((unsigned long long)w * h),
           sum[2] / ((unsigned long long)w * h), sum[3] / ((unsigned long long)w * h));

    /* encode_BMP wants room for the 54 header bytes, the pixels and one spare byte (decode_bmp.c:303-311) */
    const int64_t cap = (int64_t)(54 + rgba_size + 1);
    char *out = (char *)malloc((size_t)cap);
    uint32_t out_size = 0;
    encode_BMP(rgba, rgba_size, w, h, out, &out_size, cap);
    printf("encode_BMP wrote: %u bytes\n", out_size);
    if (out_size == 0) return 1;
    uint8_t *again = (uint8_t *)malloc(rgba_size);
    decode_BMP((const uint8_t *)out, out_size, again, (int64_t)rgba_size, &good);
    printf("round trip: %s\n", good && memcmp(again, rgba, rgba_size) == 0 ? "EXACT" : "MISMATCH");
    if (argc == 3) {
        FILE *o = fopen(argv[2], "wb");
        if (o) {
            fwrite(out, 1, out_size - 1, o);
            fclose(o);
        }
    }
    free(again);
    free(out);
    free(rgba);
    free(buf);
    return 0;
}
