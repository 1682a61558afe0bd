/*
 * fixed_deflate.c -- corpus tool (host C, not on the decode path): a small LZ77 +
 * fixed-Huffman DEFLATE *encoder* that emits ONE final block (BFINAL=1, BTYPE=01),
 * i.e. the stream shape stb_image_write produces for PNG data (BASELINE configs 3-4
 * are "via stb_write"). Written for this repository so that bench.py can make its
 * synthetic PNG corpus without touching the oracle; greedy hash-chain matcher,
 * matches of 3..258 bytes at distances up to 32767.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint8_t *p; size_t n, cap; uint64_t acc; int bits; } BitOut;

static void put(BitOut *b, uint32_t v, int n)  /* LSB first */
{
    b->acc |= (uint64_t)v << b->bits;
    b->bits += n;
    while (b->bits >= 8) {
        if (b->n < b->cap) b->p[b->n] = (uint8_t)b->acc;
        b->n++;
        b->acc >>= 8;
        b->bits -= 8;
    }
}
static uint32_t rev(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}
static void put_code(BitOut *b, uint32_t code, int n) { put(b, rev(code, n), n); }  /* Huffman codes go MSB first */

static void put_litlen(BitOut *b, uint32_t s)
{
    if (s < 144) put_code(b, 0x30 + s, 8);
    else if (s < 256) put_code(b, 0x190 + (s - 144), 9);
    else if (s < 280) put_code(b, s - 256, 7);
    else put_code(b, 0xC0 + (s - 280), 8);
}

static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LXB[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DXB[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

#define HBITS 15
#define CHAIN 8

/* Returns the number of bytes the stream needs; it was written completely iff that is <= cap. */
size_t dbg_fixed_deflate(const uint8_t *in, size_t n, uint8_t *out, size_t cap)
{
    BitOut b = {out, 0, cap, 0, 0};
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) << HBITS);
    int32_t *prev = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    for (size_t i = 0; i < ((size_t)1 << HBITS); i++) head[i] = -1;
    put(&b, 1, 1);  /* BFINAL */
    put(&b, 1, 2);  /* BTYPE = 01 fixed */
    size_t i = 0;
    while (i < n) {
        uint32_t best = 0, bdist = 0;
        if (i + 3 <= n) {
            uint32_t h = ((uint32_t)in[i] * 2654435761u ^ (uint32_t)in[i + 1] * 40503u ^ (uint32_t)in[i + 2] * 2246822519u) >> (32 - HBITS);
            int32_t c = head[h];
            for (int k = 0; k < CHAIN && c >= 0 && i - (size_t)c <= 32767; k++, c = prev[c]) {
                uint32_t l = 0, mx = n - i < 258 ? (uint32_t)(n - i) : 258;
                while (l < mx && in[c + l] == in[i + l]) l++;
                if (l > best) { best = l; bdist = (uint32_t)(i - (size_t)c); }
            }
            prev[i] = head[h];
            head[h] = (int32_t)i;
        }
        if (best >= 3) {
            int ls = 28;
            while (LBASE[ls] > best) ls--;
            put_litlen(&b, 257 + ls);
            put(&b, best - LBASE[ls], LXB[ls]);
            int ds = 29;
            while (DBASE[ds] > bdist) ds--;
            put_code(&b, ds, 5);
            put(&b, bdist - DBASE[ds], DXB[ds]);
            for (uint32_t k = 1; k < best && i + k + 3 <= n; k++) {  /* keep the hash chains warm */
                size_t j = i + k;
                uint32_t h = ((uint32_t)in[j] * 2654435761u ^ (uint32_t)in[j + 1] * 40503u ^ (uint32_t)in[j + 2] * 2246822519u) >> (32 - HBITS);
                prev[j] = head[h];
                head[h] = (int32_t)j;
            }
            i += best;
        } else {
            put_litlen(&b, in[i]);
            i++;
        }
    }
    put_litlen(&b, 256);
    if (b.bits) put(&b, 0, 8 - b.bits);
    free(head);
    free(prev);
    return b.n;
}
