/*
 * oracle/debig_oracle.c -- CPU restatement of the reference's inflate / PNG /
 * gzip decode path. TEST INFRASTRUCTURE ONLY: only tests/, smoke() and
 * bench.py's CPU legs may load it; the product never does.
 *
 * It restates WHAT the reference computes (silent, no-assert build), in plain
 * C99 written for this repository, so that the checker can travel to machines
 * where /root/reference does not exist. It is pinned: tests/test_oracle.py
 * checks it against every golden vector in tests/golden/manifest.json, which
 * was produced by the unmodified reference (oracle/_ref/libref.so), and, when
 * libref.so is present, against the reference itself on seeded random inputs.
 *
 * Deliberate, documented differences from the reference binary:
 *   D1  decode_png's output/scratch aliasing (decode_png.c:719-721,808-812)
 *       corrupts up to 771 trailing stream bytes when a late deflate block
 *       starts; that is a memory-layout accident and is NOT reproduced here.
 *   UB  where the reference has undefined behaviour (output overflow in silent
 *       builds, reads past the input, litlen symbols 286/287) this code fails
 *       the stream.
 */
#include "debig_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------ bit reader ---- */
/* LSB-first reader, inflate.c:84-91,225-278,367-413. `bitpos` counts consumed
 * bits; the reference's byte cursor is ceil(bitpos / 8). */
typedef struct {
    const uint8_t *p;
    uint64_t size;   /* declared compressed size */
    uint64_t avail;  /* bytes that may actually be read (size + padding) */
    uint64_t bitpos;
} BitIn;

static uint32_t peek(const BitIn *b, uint32_t n)
{
    uint64_t byte = b->bitpos >> 3;
    uint32_t sh = (uint32_t)(b->bitpos & 7);
    uint64_t acc = 0;
    for (uint32_t k = 0; k < 6; k++) {
        uint64_t v = (byte + k < b->avail) ? b->p[byte + k] : 0;
        acc |= v << (8 * k);
    }
    acc >>= sh;
    return (uint32_t)(acc & ((n >= 32) ? 0xffffffffull : ((1ull << n) - 1)));
}
static uint32_t take(BitIn *b, uint32_t n)
{
    uint32_t v = peek(b, n);
    b->bitpos += n;
    return v;
}

static uint32_t rev_bits(uint32_t v, uint32_t n) /* inflate.c:151-220 */
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

/* ----------------------------------------------------------- Huffman code ---- */
/* unpack_huffman (inflate.c:565-706) + huffman_to_hashmap (:494-557): canonical
 * codes, a symbol is usable iff its length is non-zero; decoding probes lengths
 * min..max in ascending order (:437-463), where "max" is the reference's own
 * bookkeeping value (Q5, :528-539). */
typedef struct {
    uint16_t first[16];  /* first canonical code of each length */
    uint16_t count[16];
    uint16_t offs[16];
    uint16_t sorted[288];
    uint32_t min_len, max_len; /* as the reference tracks them: 9999 / 1 initially */
} Huff;

static int huff_build(Huff *h, const uint32_t *lens, uint32_t n)
{
    uint32_t bl[19];
    memset(bl, 0, sizeof bl);
    memset(h, 0, sizeof *h);
    uint32_t maxl = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (lens[i] >= n) return 0; /* inflate.c:599-602 (Q3) */
        if (lens[i] > 15) return 0;
        if (lens[i] > maxl) maxl = lens[i];
        bl[lens[i]]++;
    }
    bl[0] = 0;
    uint32_t code = 0, off = 0;
    for (uint32_t bits = 1; bits <= 15; bits++) { /* :636-642 */
        code = (code + bl[bits - 1]) << 1;
        if (bl[bits] && code + bl[bits] > (1u << bits)) return 0; /* over-subscribed: not reproduced */
        h->first[bits] = (uint16_t)code;
        h->count[bits] = (uint16_t)bl[bits];
        h->offs[bits] = (uint16_t)off;
        off += bl[bits];
    }
    uint32_t fill[16];
    for (uint32_t b = 0; b < 16; b++) fill[b] = h->offs[b];
    h->min_len = 9999; /* construct_hashed_huffman :485-486 */
    h->max_len = 1;
    for (uint32_t s = 0; s < n; s++) {
        uint32_t l = lens[s];
        if (!l) continue;
        h->sorted[fill[l]++] = (uint16_t)s;
        if (l < h->min_len) h->min_len = l;     /* :528-539, the `else if` is the quirk */
        else if (l > h->max_len) h->max_len = l;
    }
    return 1;
}

/* hashed_huffman_decode inflate.c:421-474. Returns the symbol or -1. */
static int huff_decode(const Huff *h, BitIn *b)
{
    uint32_t bits = peek(b, 24);
    for (uint32_t l = h->min_len; l <= h->max_len && l <= 15; l++) {
        uint32_t c = rev_bits(bits & ((1u << l) - 1), l);
        uint32_t idx = c - h->first[l];
        if (c >= h->first[l] && idx < h->count[l]) {
            b->bitpos += l;
            return h->sorted[h->offs[l] + idx];
        }
    }
    return -1;
}

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59,
                                      67, 83, 99, 115, 131, 163, 195, 227, 258}; /* inflate.c:716-746 */
static const uint8_t LEN_XB[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                       1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577}; /* :748-779 */
static const uint8_t DIST_XB[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15}; /* :25-26 */

/* symbol statistics of the last oracle_inflate() call (for workload characterisation in DESIGN.md) */
static uint64_t g_stats[8]; /* literals, matches, match bytes, dist<=32, dist>4096, len<=8, blocks, stored bytes */
void oracle_stats(uint64_t *out8) { memcpy(out8, g_stats, sizeof g_stats); }

/* ---------------------------------------------------------------- inflate ---- */
void oracle_inflate(const uint8_t *in, uint64_t in_size, uint64_t in_avail, uint8_t *out, uint64_t cap,
                    uint64_t *out_size, uint32_t *good)
{
    *good = 0;
    *out_size = 0;
    if (!in || !out) return;
    if (cap < in_size) return;  /* inflate.c:826 */
    if (in_size < 5) return;    /* :836 */
    BitIn b = {in, in_size, in_avail < in_size ? in_size : in_avail, 0};
    memset(g_stats, 0, sizeof g_stats);
    uint64_t pos = 0;
    int more = 1;
    static Huff lit, dist, cl; /* single-threaded checker */
    uint32_t lens[320 + 140];
    while (more) {
        if (b.bitpos >= 8 * in_size) return; /* the reference would read past the input here */
        uint32_t bfinal = take(&b, 1);       /* :901-917 */
        uint32_t btype = take(&b, 2);
        if (bfinal) more = 0;
        g_stats[6]++;
        if (btype == 0) { /* :919-989 */
            b.bitpos = (b.bitpos + 7) & ~7ull;
            uint32_t len = take(&b, 16), nlen = take(&b, 16);
            if (len != ((~nlen) & 0xffff)) return; /* :949 */
            uint64_t at = b.bitpos >> 3;
            if (at + len > in_size) return;
            if (pos + len > cap) return;
            memcpy(out + pos, in + at, len);
            g_stats[7] += len;
            pos += len;
            b.bitpos += 8ull * len;
            continue;
        }
        if (btype == 3) continue; /* :990-998 */
        uint32_t hlit = 288, hdist = 0;
        int fixed = btype == 1;
        if (fixed) { /* :1035-1084 */
            for (uint32_t i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        } else { /* :1204-1520 */
            hlit = take(&b, 5) + 257;
            hdist = take(&b, 5) + 1;
            uint32_t hclen = take(&b, 4) + 4;
            uint32_t cll[19];
            memset(cll, 0, sizeof cll);
            for (uint32_t i = 0; i < hclen; i++) cll[CL_ORDER[i]] = take(&b, 3);
            if (!huff_build(&cl, cll, 19)) return;
            uint32_t n = hlit + hdist, i = 0;
            while (i < n) {
                int s = huff_decode(&cl, &b);
                if (s < 0) return;
                if (s <= 15) {
                    lens[i++] = (uint32_t)s;
                } else if (s == 16) {
                    if (i == 0) return; /* the reference reads table[-1] */
                    uint32_t rep = take(&b, 2) + 3, prev = lens[i - 1];
                    for (uint32_t k = 0; k < rep; k++) lens[i + k] = prev;
                    i += rep;
                } else if (s == 17) {
                    uint32_t rep = take(&b, 3) + 3;
                    for (uint32_t k = 0; k < rep; k++) lens[i + k] = 0;
                    i += rep;
                } else {
                    uint32_t rep = take(&b, 7) + 11;
                    for (uint32_t k = 0; k < rep; k++) lens[i + k] = 0;
                    i += rep;
                }
            }
        }
        if (!huff_build(&lit, lens, hlit)) return;                 /* :1543-1558 */
        if (!fixed && !huff_build(&dist, lens + hlit, hdist)) return; /* :1615-1630 */
        for (;;) { /* :1697-1909 */
            if (((b.bitpos + 7) >> 3) >= in_size) { /* Q2, :1702-1717 */
                more = 0;
                break;
            }
            int s = huff_decode(&lit, &b);
            if (s < 0) return;
            if (s < 256) {
                if (pos >= cap) return;
                out[pos++] = (uint8_t)s;
                g_stats[0]++;
                continue;
            }
            if (s == 256) break;
            if (s > 285) return;
            uint32_t len = LEN_BASE[s - 257] + (LEN_XB[s - 257] ? take(&b, LEN_XB[s - 257]) : 0);
            uint32_t ds;
            if (fixed) {
                ds = rev_bits(take(&b, 5), 5); /* :1783-1788 */
            } else {
                int d = huff_decode(&dist, &b);
                if (d < 0) return;
                ds = (uint32_t)d;
            }
            if (ds > 29) return; /* :1809 */
            uint32_t dd = DIST_BASE[ds] + (DIST_XB[ds] ? take(&b, DIST_XB[ds]) : 0);
            if (dd > pos) return; /* :1843 */
            if (pos + len > cap) return;
            g_stats[1]++; g_stats[2] += len; g_stats[3] += dd <= 32; g_stats[4] += dd > 4096; g_stats[5] += len <= 8;
            for (uint32_t k = 0; k < len; k++) out[pos + k] = out[pos + k - dd]; /* :1861-1897 */
            pos += len;
        }
    }
    *out_size = pos;
    *good = 1;
}

/* ------------------------------------------------------------------- gzip ---- */
/* decode_gz.c:123-233 (silent build) + :270. */
void oracle_decode_gz(const uint8_t *in, uint64_t in_size, uint8_t *out, uint64_t cap, uint64_t *out_size, uint32_t *good)
{
    *good = 0;
    *out_size = 0;
    if (!in || in_size < 10) return;
    if (in[0] != 31 || in[1] != 139 || in[2] != 8) return;
    uint64_t at = 10, left = in_size - 10;
    if ((in[3] >> 3) & 1) {
        uint64_t n = 0;
        while (n < left && in[at + n] != 0) n++;
        if (n + 1 > left) return;
        at += n + 1;
        left -= n + 1;
    }
    if (left < 8) return;
    oracle_inflate(in + at, left - 8, left, out, cap, out_size, good);
}

/* -------------------------------------------------------------------- PNG ---- */
static uint32_t crc_tab[256];
static void crc_init(void) /* decode_png.c:289-304 */
{
    if (crc_tab[1]) return;
    for (uint32_t n = 0; n < 256; n++) {
        uint32_t c = n;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        crc_tab[n] = c;
    }
}
static uint32_t crc_run(uint32_t c, const uint8_t *p, uint64_t n) /* :313-333 */
{
    for (uint64_t i = 0; i < n; i++) c = crc_tab[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c;
}
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static uint8_t paeth(int a, int b, int c) /* :441-487 */
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (uint8_t)((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c));
}
static uint8_t unfilter(uint32_t ft, uint8_t x, uint8_t a, uint8_t b, uint8_t c) /* :497-541 */
{
    switch (ft) {
        case 0: return x;
        case 1: return (uint8_t)(x + a);
        case 2: return (uint8_t)(x + b);
        case 3: return (uint8_t)(x + (uint8_t)(((uint32_t)a + b) / 2));
        case 4: return (uint8_t)(x + paeth(a, b, c));
        default: return 0;
    }
}

void oracle_png_dims(const uint8_t *in, uint64_t in_size, uint32_t *w, uint32_t *h, uint8_t *good)
{
    *w = *h = 0;
    *good = 0;
    if (in_size < 28) return;               /* decode_png.c:627 */
    if (memcmp(in + 1, "PNG", 3)) return;   /* :644 */
    *w = be32(in + 16);
    *h = be32(in + 20);
    *good = 1;
}

/* decode_png.c:683-1567. `rgb_as_reference` != 0 reproduces the reference's
 * per-row 3->4 expansion (D3, :1509-1536); 0 expands correctly once at the end. */
void oracle_decode_png(const uint8_t *file, uint64_t size, uint8_t *out, uint64_t rgba_size, int rgb_as_reference,
                       uint8_t *good)
{
    *good = 0;
    crc_init();
    if (size < 8 || memcmp(file + 1, "PNG", 3)) return;
    uint64_t pos = 8, left = size - 8;
    int found_idat = 0, ran = 0, found_ihdr = 0, found_iend = 0;
    uint32_t w = 0, h = 0, ct = 0;
    uint8_t pal[768];
    uint32_t pal_n = 0;
    memset(pal, 0, sizeof pal);
    uint8_t *z = (uint8_t *)malloc(size + 64);
    uint64_t zlen = 0, zrun = 0;
    uint8_t *scan = NULL;
    if (!z) return;
    memset(z, 0, size + 64);
    while (left >= 8 && !found_iend) { /* :755 */
        if (pos + 8 > size) goto done;
        uint64_t length = be32(file + pos);
        const uint8_t *type = file + pos + 4;
        pos += 8;
        left -= 8;
        int is_idat = !memcmp(type, "IDAT", 4);
        if (!is_idat && found_idat) { /* :775: inflate runs here */
            ran = 1;
            zrun = zlen;
        }
        if (length >= left) goto done;          /* :886 */
        if (pos + length + 4 > size) goto done; /* real bounds */
        uint32_t crc = crc_run(0xffffffffu, type, 4 + length) ^ 0xffffffffu; /* :862-874 */
        if (!memcmp(type, "PLTE", 4)) { /* :900-950 */
            if (!found_ihdr) goto done;
            if (length % 3) goto done;
            pal_n = (uint32_t)(length / 3);
            for (uint32_t i = 0; i < pal_n && i < 256; i++) memcpy(pal + 3 * i, file + pos + 3 * i, 3);
            pos += length;
        } else if (!memcmp(type, "IHDR", 4)) { /* :951-1138 */
            found_ihdr = 1;
            if (pos + 13 > size) goto done;
            const uint8_t *b = file + pos;
            pos += 13;
            w = be32(b);
            h = be32(b + 4);
            ct = b[9];
            if ((uint64_t)w * h * 4 != rgba_size) goto done;
            if ((uint64_t)w * h * 4 + h + 1 >= (1ull << 32)) goto done;
            if (!(ct == 2 || ct == 3 || ct == 6)) goto done;
            if (w < 1 || h < 1) goto done;
            if (b[8] != 8) goto done;
            if (b[11] != 0) goto done;
            if (left < 4) goto done;
        } else if (is_idat) { /* :1140-1292 */
            if (!found_ihdr) goto done;
            uint64_t dl = length;
            if (!found_idat) {
                found_idat = 1;
                if (length < 2) goto done;
                uint32_t cmf = file[pos], flg = file[pos + 1];
                pos += 2;
                dl -= 2;
                if ((cmf & 15) != 8) goto done;
                uint32_t chk = (cmf << 8) | flg;
                if (chk == 0 || chk % 31) goto done;
                if ((flg >> 5) & 1) goto done;
            }
            memcpy(z + zlen, file + pos, dl); /* :1285-1291 */
            zlen += dl;
            pos += dl;
            left -= dl;
        } else if (!memcmp(type, "IEND", 4)) {
            found_iend = 1;
        } else if ((signed char)type[0] > 'Z') { /* :1303 */
            pos += length;
            left -= length;
        } else {
            goto done;
        }
        if (left < 4) goto done;
        if (pos + 4 > size) goto done;
        uint32_t stored = be32(file + pos);
        pos += 4;
        left -= 4;
        if (stored != crc) goto done; /* :1341-1348 */
    }
    if (!ran) goto done; /* :1357 */
    if (zrun < 4) goto done;
    {
        uint64_t est = (uint64_t)w * h * 4 + h + 1, got = 0; /* :965-968 */
        uint32_t ok = 0;
        scan = (uint8_t *)malloc(est + 64);
        if (!scan) goto done;
        memset(scan, 0, est + 64);
        oracle_inflate(z, zrun - 4, zlen + 32, scan, est, &got, &ok); /* :800-820 */
        if (!ok) goto done;
        if (scan[0] > 4) goto done; /* :847-858 */
        uint32_t bpp = ct == 6 ? 4 : ct == 2 ? 3 : 1; /* :1401-1414 */
        uint64_t stride = (uint64_t)w * bpp;
        if (got < (uint64_t)h * (stride + 1)) goto done; /* Q13: the reference reads stale memory */
        uint64_t npx = (uint64_t)w * h;
        /* :1430-1507: the reference writes un-filtered bytes contiguously into out */
        for (uint32_t r = 0; r < h; r++) {
            const uint8_t *row = scan + (uint64_t)r * (stride + 1);
            uint32_t ft = row[0];
            uint8_t *o = out + (uint64_t)r * stride;
            for (uint64_t i = 0; i < stride; i++) {
                uint8_t a = i >= bpp ? o[i - bpp] : 0;
                uint8_t b = r > 0 ? o[(int64_t)i - (int64_t)stride] : 0;
                uint8_t c = (r > 0 && i >= bpp) ? o[(int64_t)i - (int64_t)stride - bpp] : 0;
                o[i] = unfilter(ft, row[1 + i], a, b, c);
            }
            if (bpp == 3 && rgb_as_reference) { /* :1509-1536, executed once per row */
                uint8_t *wr = out + npx * 4 - 1, *rd = out + npx * 3 - 1;
                for (uint64_t p = 0; p < npx; p++) {
                    *wr-- = 255;
                    *wr-- = *rd--;
                    *wr-- = *rd--;
                    *wr-- = *rd--;
                }
            }
        }
        if (bpp == 3 && !rgb_as_reference) {
            for (uint64_t p = npx; p-- > 0;) {
                uint8_t r_ = out[3 * p], g_ = out[3 * p + 1], b_ = out[3 * p + 2];
                out[4 * p] = r_;
                out[4 * p + 1] = g_;
                out[4 * p + 2] = b_;
                out[4 * p + 3] = 255;
            }
        }
        if (bpp == 1) { /* :1538-1564 */
            for (uint64_t p = npx; p-- > 0;) {
                uint32_t idx = out[p];
                out[4 * p] = idx < pal_n ? pal[3 * idx] : 0;
                out[4 * p + 1] = idx < pal_n ? pal[3 * idx + 1] : 0;
                out[4 * p + 2] = idx < pal_n ? pal[3 * idx + 2] : 0;
                out[4 * p + 3] = 255;
            }
        }
        *good = 1;
    }
done:
    free(z);
    free(scan);
}

/* ------------------------------------------------------------------- BMP --
 * decode_bmp.c:53-103 (get_BMP_width_height), :105-295 (decode_BMP), :297-372
 * (encode_BMP), release semantics (asserts off). Where the reference reads or
 * writes out of bounds (undefined behaviour) this restatement reports good = 0,
 * exactly like the product; those inputs are outside the parity domain. */
static uint32_t bmp_u32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t bmp_u16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

void oracle_bmp_dims(const uint8_t *in, uint64_t size, uint32_t *w, uint32_t *h, uint8_t *good)
{
    *good = 0;
    if (size < 54) return;                      /* asserted in the reference (:69-80) */
    if (in[0] != 'B' || in[1] != 'M') return;   /* :84-99 */
    int32_t width = (int32_t)bmp_u32(in + 18), height = (int32_t)bmp_u32(in + 22);
    *h = (uint32_t)(height < 0 ? -(int64_t)height : height); /* :101-105 */
    *w = (uint32_t)width;
    *good = *w > 0 && *h > 0;
}

void oracle_decode_bmp(const uint8_t *in, uint64_t size, uint8_t *out, int64_t out_size, uint8_t *good)
{
    *good = 0;
    if (size < 54 || out_size < 0) return;
    if (in[0] != 'B' || in[1] != 'M') return;                                /* :120-133 */
    uint32_t bf_size = bmp_u32(in + 2), image_offset = bmp_u32(in + 10);
    if ((uint64_t)(uint32_t)(image_offset + bf_size) + 14 < size) return;    /* :135-150: file longer than declared */
    uint32_t dib = bmp_u32(in + 14);
    if (dib != 40 && dib != 108) return;                                     /* :158-178 */
    int32_t width = (int32_t)bmp_u32(in + 18), height = (int32_t)bmp_u32(in + 22);
    int bottom_first = 1;
    if (height < 0) {                                                        /* :180-186 */
        bottom_first = 0;
        if (height == INT32_MIN) return;
        height = -height;
    }
    /* :188-202 compares w*h*4 with out_size but the verdict is overwritten by :293 */
    if (bmp_u16(in + 26) != 1) return;                                       /* :204-211 */
    if (bmp_u16(in + 28) != 32) return;                                      /* :213-221 */
    /* :223-258 compression / resolution / palette counts: no effect (see :293) */
    if (width < 0) return;                                                   /* UB domain */
    uint64_t need = (uint64_t)(uint32_t)width * (uint32_t)height * 4;
    if (need > 0xffffffffull || need > (uint64_t)out_size || (uint64_t)image_offset + need > size) return; /* UB domain */
    uint32_t w = (uint32_t)width, h = (uint32_t)height;
    for (uint32_t y = 0; y < h; y++) {                                       /* :260-291 */
        uint32_t ty = bottom_first ? h - y - 1 : y;
        const uint8_t *s = in + image_offset + (uint64_t)y * 4 * w;
        uint8_t *d = out + (uint64_t)ty * 4 * w;
        for (uint32_t x = 0; x < w; x++) {
            d[4 * x + 0] = s[4 * x + 2];
            d[4 * x + 1] = s[4 * x + 1];
            d[4 * x + 2] = s[4 * x + 0];
            d[4 * x + 3] = s[4 * x + 3];
        }
    }
    *good = 1;                                                               /* :293 */
}

void oracle_encode_bmp(const uint8_t *rgba, uint64_t rgba_size, uint32_t w, uint32_t h, uint8_t *out,
                       uint32_t *out_size, int64_t cap)
{
    *out_size = 0;
    uint64_t total = 14 + 40 + rgba_size + 1;                                /* :311 */
    if ((rgba_size & 3) || cap < 0 || total > (uint64_t)cap || total > 0xffffffffull) return; /* UB domain */
    *out_size = (uint32_t)total;
    uint8_t hd[54];
    memset(hd, 0, sizeof hd);
#define PUT32(at, v) do { uint32_t v_ = (v); hd[at] = (uint8_t)v_; hd[(at) + 1] = (uint8_t)(v_ >> 8); hd[(at) + 2] = (uint8_t)(v_ >> 16); hd[(at) + 3] = (uint8_t)(v_ >> 24); } while (0)
    hd[0] = 'B'; hd[1] = 'M';                                                /* :315-326 */
    PUT32(2, w * h * 4 + 54);
    PUT32(10, 54);
    PUT32(14, 40);                                                           /* :333-347 */
    PUT32(18, w);
    PUT32(22, (uint32_t)(-(int32_t)h));
    hd[26] = 1;
    hd[28] = 32;
    PUT32(34, w * h * 4);
#undef PUT32
    memcpy(out, hd, 54);
    for (uint64_t i = 0; i < rgba_size; i += 4) {                            /* :354-364 */
        out[54 + i + 0] = rgba[i + 2];
        out[54 + i + 1] = rgba[i + 1];
        out[54 + i + 2] = rgba[i + 0];
        out[54 + i + 3] = rgba[i + 3];
    }
}
