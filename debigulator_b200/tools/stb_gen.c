/*
 * stb_gen.c -- CORPUS TOOLING, not product code: synthetic PNG / zlib writer for benchmarks and tests.
 *
 * BASELINE.json names "the repo's bundled stb_write.h PNG/zlib writer" as the generator of the synthetic PNG corpora
 * (configs 3 and 4). That file is the reference's vendored stb_image_write v1.16 (src/stb_write.h). It is NOT copied
 * into this repository: debigulator_b200/build.py compiles this shim against it where it lies
 * (-I/root/reference/src) into debigulator_b200/tools/libstbgen.so, which is git-ignored and travels to the GPU box
 * with the working tree like the other built libraries. Without the reference tree and without a prebuilt library
 * the corpus falls back to tools/fixed_deflate.c (same stream shape) and says so.
 *
 *   stbgen_png   stbi_write_png_to_mem (stb_write.h:1128) with stbi_write_force_png_filter (:253): filter -1 =
 *                per-row adaptive, 0..4 = None / Sub / Up / Avg / Paeth forced
 *   stbgen_zlib  stbi_zlib_compress (stb_write.h:895): zlib header 78 5E, ONE final fixed-Huffman block, Adler-32
 */
#include <stdint.h>
#include <stdlib.h>

#define STB_IMAGE_WRITE_IMPLEMENTATION
#define STBI_WRITE_NO_STDIO
#define STB_IMAGE_WRITE_STATIC
#include "stb_write.h"

uint8_t *stbgen_png(const uint8_t *pixels, int w, int h, int comp, int filter, int *out_len)
{
    stbi_write_force_png_filter = filter;
    return stbi_write_png_to_mem(pixels, w * comp, w, h, comp, out_len);
}

uint8_t *stbgen_zlib(uint8_t *data, int len, int quality, int *out_len) { return stbi_zlib_compress(data, len, out_len, quality); }

void stbgen_free(void *p) { free(p); }
