/* tests/tools/sync_probe.c -- measurement tool (not product code): how quickly does a DEFLATE
 * symbol decoder that starts at an arbitrary bit offset inside a block fall into step with the
 * true symbol sequence? Used to size the lead-in of the sub-chunk parallel decoder.
 * Reuses the oracle port's Huffman helpers by including its source. */
#include "../../oracle/debig_oracle.c"
#include <stdio.h>

static uint8_t *starts; /* bitmap of true symbol starts */

/* decode one symbol at b (tables lit/dist given); returns 0 ok, 1 eob, -1 error */
static int one_symbol(BitIn *b, const Huff *lit, const Huff *dist, int fixed)
{
    int s = huff_decode(lit, b);
    if (s < 0) return -1;
    if (s < 256) return 0;
    if (s == 256) return 1;
    if (s > 285) return -1;
    if (LEN_XB[s - 257]) take(b, LEN_XB[s - 257]);
    uint32_t ds;
    if (fixed) ds = rev_bits(take(b, 5), 5);
    else { int d = huff_decode(dist, b); if (d < 0) return -1; ds = (uint32_t)d; }
    if (ds > 29) return -1;
    if (DIST_XB[ds]) take(b, DIST_XB[ds]);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 2) return 1;
    FILE *f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *in = malloc(n + 64); memset(in, 0, n + 64);
    if (fread(in, 1, n, f) != (size_t)n) return 1;
    int SUB = argc > 2 ? atoi(argv[2]) : 288;
    starts = calloc(n + 64, 1);
    BitIn b = {in, (uint64_t)n, (uint64_t)n + 32, 0};
    static Huff lit, dist, cl; uint32_t lens[460];
    int leads[] = {32, 64, 128, 192, 256, 384, 512, 1024};
    long tries[8] = {0}, ok[8] = {0};
    int more = 1;
    while (more) {
        uint32_t bfinal = take(&b, 1), btype = take(&b, 2);
        if (bfinal) more = 0;
        if (btype == 0) { b.bitpos = (b.bitpos + 7) & ~7ull; uint32_t len = take(&b, 16); take(&b, 16); b.bitpos += 8ull * len; continue; }
        if (btype == 3) continue;
        uint32_t hlit = 288, hdist = 0; int fixed = btype == 1;
        if (fixed) { for (uint32_t i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8; }
        else {
            hlit = take(&b, 5) + 257; hdist = take(&b, 5) + 1; uint32_t hclen = take(&b, 4) + 4; uint32_t cll[19]; memset(cll, 0, sizeof cll);
            for (uint32_t i = 0; i < hclen; i++) cll[CL_ORDER[i]] = take(&b, 3);
            huff_build(&cl, cll, 19);
            uint32_t nn = hlit + hdist, i = 0;
            while (i < nn) { int s = huff_decode(&cl, &b);
                if (s <= 15) lens[i++] = s;
                else if (s == 16) { uint32_t rep = take(&b, 2) + 3, prev = lens[i - 1]; for (uint32_t k = 0; k < rep; k++) lens[i + k] = prev; i += rep; }
                else if (s == 17) { uint32_t rep = take(&b, 3) + 3; for (uint32_t k = 0; k < rep; k++) lens[i + k] = 0; i += rep; }
                else { uint32_t rep = take(&b, 7) + 11; for (uint32_t k = 0; k < rep; k++) lens[i + k] = 0; i += rep; } }
        }
        huff_build(&lit, lens, hlit);
        if (!fixed) huff_build(&dist, lens + hlit, hdist);
        uint64_t blk_start = b.bitpos;
        for (;;) { starts[b.bitpos >> 3] |= 1 << (b.bitpos & 7); int r = one_symbol(&b, &lit, &dist, fixed); if (r) break; }
        uint64_t blk_end = b.bitpos;
        /* probe boundaries inside this block */
        for (uint64_t B = blk_start + 2048; B + 64 < blk_end; B += (uint64_t)SUB * 7) {
            uint64_t truth = B; while (!(starts[truth >> 3] & (1 << (truth & 7)))) truth++;
            for (int li = 0; li < 8; li++) {
                BitIn p = {in, (uint64_t)n, (uint64_t)n + 32, B - leads[li]};
                int r = 0; while (p.bitpos < B && r == 0) r = one_symbol(&p, &lit, &dist, fixed);
                tries[li]++; if (r == 0 && p.bitpos == truth) ok[li]++;
            }
        }
    }
    for (int li = 0; li < 8; li++) printf("lead %4d bits: synced %.4f (%ld probes)\n", leads[li], tries[li] ? (double)ok[li] / tries[li] : 0, tries[li]);
    return 0;
}
