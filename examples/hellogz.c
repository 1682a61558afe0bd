/*
 * examples/hellogz.c -- a fresh example main in the role of the reference's src/hellogz.c
 * (which does not compile against the current inflate.h): decode one gzip
 * file through the drop-in API of include/decode_gz.h and print its size and first bytes.
 *
 *   gcc -std=c99 -Iinclude examples/hellogz.c -Ldebigulator_b200 -ldebigulator_b200 \
 *       -Wl,-rpath,$PWD/debigulator_b200 -o hellogz
 *   ./hellogz tests/golden/gzipsample.gz
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "decode_gz.h"

int main(int argc, char **argv)
{
    if (argc != 2) {
        fprintf(stderr, "usage: %s file.gz\n", argv[0]);
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
    init_decode_gz(malloc, memset, memcpy);
    DecodedData *d = decode_gz(buf, (uint32_t)got);
    if (!d || !d->good) {
        printf("decode_gz result was: FAILURE\n");
        return 1;
    }
    printf("decode_gz result was: SUCCESS\n");
    printf("decompressed bytes: %u\n", d->data_size);
    unsigned long long sum = 0;
    for (uint32_t i = 0; i < d->data_size; i++) sum += (unsigned char)d->data[i];
    printf("byte sum: %llu\n", sum);
    printf("first line: %.*s\n", (int)(strcspn(d->data, "\n") < 100 ? strcspn(d->data, "\n") : 100), d->data);
    free(d->data);
    free(d);
    free(buf);
    return 0;
}
