// bmp_kernels.cuh -- batched BMP decode / encode.
//
// Reference: decode_bmp.c:105-295 (decode_BMP), :297-372 (encode_BMP). Both are a
// header check followed by a BGRA <-> RGBA byte swizzle, decode optionally flipping
// the row order (positive DIB height = bottom-up file). Pure HBM traffic: every pixel
// is read once and written once, so this is the one kernel of the library that is
// measured against the copy roofline rather than against instruction issue.
//
// Reference behaviour kept:
//   * only 'BM', DIB header size 40 or 108, planes == 1, 32 bits per pixel decode
//     (decode_bmp.c:120-133, :158-178, :204-221);
//   * the file is rejected when it is LONGER than image_offset + bfSize + 14
//     (decode_bmp.c:135-150, 32-bit sum);
//   * compression / resolution / palette fields and the output-size comparison only
//     set good = 0 transiently -- the function ends with good = 1 (decode_bmp.c:293),
//     so they do not reject anything;
//   * encode writes a 40-byte-DIB top-down file (negative height) and reports
//     54 + rgba_size + 1 bytes, the last of which it never writes (decode_bmp.c:311).
// Divergences (undefined behaviour in the reference, reported here instead):
//   * pixel data reaching past the input            -> ST_TRUNCATED
//   * w*h*4 larger than the output capacity         -> ST_OUT_OVERFLOW
//   * negative width, or more than 2^32 output bytes -> ST_TOO_LARGE
//   * input shorter than the 54 header bytes        -> ST_CONTAINER
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "inflate_core.h"

namespace dbg {

constexpr uint32_t BMP_TILE_PIXELS = 4096;  // 16 KiB of pixels per CTA step
constexpr int BMP_THREADS = 256;

struct BmpItem {
    uint64_t src;        // byte offset of the first pixel from the batch's source base
    uint64_t dst;        // byte offset of the first pixel from the batch's destination base
    uint32_t w, h;       // rows of w pixels; encode and top-down decode use one row of w*h... (see bmp_plan)
    uint32_t flip;       // 1 = row r of the source is row h-1-r of the destination
    uint32_t rows_per_tile, segs_per_row;
};

struct BmpBatch {
    const uint8_t *in_base;
    const uint64_t *in_off;
    const uint64_t *in_size;   // decode: file bytes; encode: rgba bytes
    uint8_t *out_base;
    const uint64_t *out_off;
    const uint64_t *out_cap;
    uint64_t *out_size;
    uint32_t *status;
    uint32_t *width, *height;  // decode: optional outputs; encode: inputs
    uint32_t n;
    // scratch
    BmpItem *items;
    uint32_t *tile_base;       // n + 1 entries after bmp_scan_kernel
};

__device__ __forceinline__ uint32_t bmp_rd32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint32_t bmp_rd16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

__device__ __forceinline__ void bmp_set_tiles(BmpItem &it, uint32_t *ntiles)
{
    // wide rows are cut into segments of BMP_TILE_PIXELS, narrow rows are grouped
    if (it.w == 0 || it.h == 0) {
        it.rows_per_tile = 1;
        it.segs_per_row = 0;
        *ntiles = 0;
        return;
    }
    it.segs_per_row = (it.w + BMP_TILE_PIXELS - 1) / BMP_TILE_PIXELS;
    it.rows_per_tile = it.segs_per_row == 1 ? (BMP_TILE_PIXELS / it.w > 0 ? BMP_TILE_PIXELS / it.w : 1) : 1;
    *ntiles = it.segs_per_row == 1 ? (it.h + it.rows_per_tile - 1) / it.rows_per_tile : it.h * it.segs_per_row;
}

// One thread per file: header checks of decode_bmp.c:105-221 and the tile count.
__global__ void bmp_decode_plan_kernel(BmpBatch b)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const uint8_t *p = b.in_base + b.in_off[i];
    const uint64_t size = b.in_size[i];
    BmpItem it = {};
    uint32_t st = ST_OK, ntiles = 0, w = 0, h = 0;
    uint64_t need = 0;
    if (size < 54 || p[0] != 'B' || p[1] != 'M') {
        st = ST_CONTAINER;
    } else {
        const uint32_t bf_size = bmp_rd32(p + 2), image_offset = bmp_rd32(p + 10), dib_size = bmp_rd32(p + 14);
        const int32_t width = (int32_t)bmp_rd32(p + 18), height = (int32_t)bmp_rd32(p + 22);
        const uint32_t planes = bmp_rd16(p + 26), bpp = bmp_rd16(p + 28);
        if ((uint64_t)(uint32_t)(image_offset + bf_size) + 14 < size) st = ST_CONTAINER;  // decode_bmp.c:135
        else if (dib_size != 40 && dib_size != 108) st = ST_CONTAINER;                    // :158-178
        else if (planes != 1 || bpp != 32) st = ST_CONTAINER;                             // :204-221
        else if (width < 0 || height == INT32_MIN) st = ST_TOO_LARGE;
        else {
            w = (uint32_t)width;
            h = (uint32_t)(height < 0 ? -height : height);
            need = (uint64_t)w * h * 4;
            if (need > 0xffffffffull) st = ST_TOO_LARGE;  // the reference indexes with 32-bit arithmetic (:262-276)
            else if (need > b.out_cap[i]) st = ST_OUT_OVERFLOW;
            else if ((uint64_t)image_offset + need > size) st = ST_TRUNCATED;
            else {
                it.src = b.in_off[i] + image_offset;
                it.dst = b.out_off[i];
                it.w = w;
                it.h = h;
                it.flip = height < 0 ? 0u : 1u;  // :180-186
                if (!it.flip) {  // top-down: one contiguous run
                    it.w = (uint32_t)(need / 4);
                    it.h = it.w ? 1 : 0;
                }
                bmp_set_tiles(it, &ntiles);
            }
        }
    }
    b.items[i] = it;
    b.tile_base[i] = ntiles;
    b.status[i] = st;
    b.out_size[i] = st == ST_OK ? need : 0;
    if (b.width) b.width[i] = st == ST_OK ? w : 0;
    if (b.height) b.height[i] = st == ST_OK ? h : 0;
}

// One thread per image: the 54 header bytes of encode_BMP (decode_bmp.c:313-358) and the tile count.
__global__ void bmp_encode_plan_kernel(BmpBatch b)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const uint64_t rgba_size = b.in_size[i];
    const uint32_t w = b.width[i], h = b.height[i];
    BmpItem it = {};
    uint32_t st = ST_OK, ntiles = 0;
    const uint64_t total = 54 + rgba_size + 1;  // :311 (the final byte is reserved for a terminator and never written)
    if (rgba_size & 3) st = ST_CONTAINER;  // the reference copies whole pixels and would read past the input
    else if (total > 0xffffffffull) st = ST_TOO_LARGE;
    else if (total > b.out_cap[i]) st = ST_OUT_OVERFLOW;
    else {
        uint8_t *o = b.out_base + b.out_off[i];
        auto w32 = [&](int at, uint32_t v) {
            o[at] = (uint8_t)v;
            o[at + 1] = (uint8_t)(v >> 8);
            o[at + 2] = (uint8_t)(v >> 16);
            o[at + 3] = (uint8_t)(v >> 24);
        };
        o[0] = 'B';
        o[1] = 'M';
        w32(2, w * h * 4 + 54);
        w32(6, 0);
        w32(10, 54);
        w32(14, 40);
        w32(18, w);
        w32(22, (uint32_t)(-(int32_t)h));
        o[26] = 1;
        o[27] = 0;
        o[28] = 32;
        o[29] = 0;
        w32(30, 0);
        w32(34, w * h * 4);
        w32(38, 0);
        w32(42, 0);
        w32(46, 0);
        w32(50, 0);
        it.src = b.in_off[i];
        it.dst = b.out_off[i] + 54;
        it.w = (uint32_t)(rgba_size / 4);
        it.h = it.w ? 1 : 0;
        it.flip = 0;
        bmp_set_tiles(it, &ntiles);
    }
    b.items[i] = it;
    b.tile_base[i] = ntiles;
    b.status[i] = st;
    b.out_size[i] = st == ST_OK ? total : 0;
}

// Exclusive scan of the per-item tile counts (in place, n + 1 entries): one CTA.
__global__ void __launch_bounds__(1024) bmp_scan_kernel(uint32_t *tile_base, uint32_t n)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? tile_base[i] : 0;
        uint32_t x = v;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= (uint32_t)d) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t s = warp_sum[threadIdx.x];
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
                if (threadIdx.x >= (uint32_t)d) s += y;
            }
            warp_sum[threadIdx.x] = s;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0);
        if (i < n) tile_base[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_base[n] = carry;
}

// One pixel: bytes 0 and 2 trade places. SA / DA = the alignment (4, 2 or 1 bytes) of the pixel addresses.
template <int SA>
__device__ __forceinline__ uint32_t bmp_load_pixel(const uint8_t *p)
{
    if (SA == 4) return __ldg(reinterpret_cast<const uint32_t *>(p));
    if (SA == 2) {
        const uint16_t *q = reinterpret_cast<const uint16_t *>(p);
        return (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 16);
    }
    return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
}
template <int DA>
__device__ __forceinline__ void bmp_store_pixel(uint8_t *p, uint32_t v)
{
    if (DA == 4) {
        __stcs(reinterpret_cast<uint32_t *>(p), v);
    } else if (DA == 2) {
        uint16_t *q = reinterpret_cast<uint16_t *>(p);
        q[0] = (uint16_t)v;
        q[1] = (uint16_t)(v >> 16);
    } else {
        p[0] = (uint8_t)v;
        p[1] = (uint8_t)(v >> 8);
        p[2] = (uint8_t)(v >> 16);
        p[3] = (uint8_t)(v >> 24);
    }
}

// Scalar fallback: one pixel per thread step, any alignment.
template <int SA, int DA>
__device__ __forceinline__ void bmp_run_scalar(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t npix,
                                               uint32_t tid, uint32_t nthreads)
{
#pragma unroll 4
    for (uint32_t j = tid; j < npix; j += nthreads)
        bmp_store_pixel<DA>(dst + (uint64_t)j * 4, __byte_perm(bmp_load_pixel<SA>(src + (uint64_t)j * 4), 0, 0x3012));
}

// Byte i of the swizzled pixel stream that starts at src (head / tail of a vector run).
__device__ __forceinline__ uint8_t bmp_stream_byte(const uint8_t *src, uint32_t i)
{
    const uint32_t c = i & 3;
    return __ldg(src + (i & ~3u) + (c == 0 ? 2u : c == 2 ? 0u : c));
}

// 16 bytes per thread step: destination chunks are 16-byte aligned, the source bytes of a chunk come
// from the two aligned 16-byte loads that cover them (every loaded chunk holds at least one byte of
// the run, so nothing outside the caller's allocation granule is touched). WO = word offset of the
// run inside the first load; `bs` = remaining byte shift. SWZ_FIRST: the source is pixel aligned
// (encode), so pixels are swizzled before the realignment; otherwise the destination is (decode)
// and they are swizzled after it.
template <int WO, bool SWZ_FIRST>
__device__ __forceinline__ void bmp_run_vec_body(const uint8_t *__restrict__ s_al, uint8_t *__restrict__ d_al, uint32_t nchunk,
                                                 uint32_t bs, uint32_t tid, uint32_t nthreads)
{
#pragma unroll 2
    for (uint32_t k = tid; k < nchunk; k += nthreads) {
        const uint4 a0 = __ldg(reinterpret_cast<const uint4 *>(s_al) + k);
        uint4 a1 = a0;
        if (WO != 0 || bs != 0) a1 = __ldg(reinterpret_cast<const uint4 *>(s_al) + k + 1);
        uint32_t x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (SWZ_FIRST) {
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = __byte_perm(x[j], 0, 0x3012);
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            o[j] = __funnelshift_r(x[WO + j], x[WO + j + 1], bs);
            if (!SWZ_FIRST) o[j] = __byte_perm(o[j], 0, 0x3012);
        }
        __stcs(reinterpret_cast<uint4 *>(d_al) + k, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// One contiguous run of npix pixels, copied with the swizzle by `nthreads` cooperating threads.
__device__ __forceinline__ void bmp_run(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, uint32_t npix, uint32_t tid,
                                        uint32_t nthreads)
{
    const uint32_t sp = (uint32_t)((uintptr_t)src & 3), dp = (uint32_t)((uintptr_t)dst & 3);
    if (sp != 0 && dp != 0) {  // neither side pixel-aligned: scalar path
        if (!(sp & 1) && !(dp & 1)) bmp_run_scalar<2, 2>(src, dst, npix, tid, nthreads);
        else if (!(sp & 1)) bmp_run_scalar<2, 1>(src, dst, npix, tid, nthreads);
        else if (!(dp & 1)) bmp_run_scalar<1, 2>(src, dst, npix, tid, nthreads);
        else bmp_run_scalar<1, 1>(src, dst, npix, tid, nthreads);
        return;
    }
    const uint32_t nbytes = npix * 4;
    uint32_t hd = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (hd > nbytes) hd = nbytes;
    const uint32_t nchunk = (nbytes - hd) >> 4, tail0 = hd + (nchunk << 4);
    // head and tail bytes
    for (uint32_t i = tid; i < hd + (nbytes - tail0); i += nthreads) {
        const uint32_t at = i < hd ? i : tail0 + (i - hd);
        dst[at] = bmp_stream_byte(src, at);
    }
    if (nchunk == 0) return;
    const uint8_t *s = src + hd;
    const uint32_t sm = (uint32_t)((uintptr_t)s & 15), bs = (sm & 3) * 8;
    const uint8_t *s_al = s - sm;
    uint8_t *d_al = dst + hd;
    if (sp == 0) {
        switch (sm >> 2) {
            case 0: bmp_run_vec_body<0, true>(s_al, d_al, nchunk, bs, tid, nthreads); break;
            case 1: bmp_run_vec_body<1, true>(s_al, d_al, nchunk, bs, tid, nthreads); break;
            case 2: bmp_run_vec_body<2, true>(s_al, d_al, nchunk, bs, tid, nthreads); break;
            default: bmp_run_vec_body<3, true>(s_al, d_al, nchunk, bs, tid, nthreads); break;
        }
    } else {
        switch (sm >> 2) {
            case 0: bmp_run_vec_body<0, false>(s_al, d_al, nchunk, bs, tid, nthreads); break;
            case 1: bmp_run_vec_body<1, false>(s_al, d_al, nchunk, bs, tid, nthreads); break;
            case 2: bmp_run_vec_body<2, false>(s_al, d_al, nchunk, bs, tid, nthreads); break;
            default: bmp_run_vec_body<3, false>(s_al, d_al, nchunk, bs, tid, nthreads); break;
        }
    }
}

// The swizzle: CTAs stride over the tiles of the whole batch.
__global__ void __launch_bounds__(BMP_THREADS) bmp_swizzle_kernel(BmpBatch b)
{
    const uint32_t total = b.tile_base[b.n];
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
        // tile -> item: last i with tile_base[i] <= t
        uint32_t lo = 0, hi = b.n;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (b.tile_base[mid] <= t) lo = mid;
            else hi = mid;
        }
        const BmpItem it = b.items[lo];
        const uint32_t k = t - b.tile_base[lo];
        uint32_t row0, nrows, x0, npix;
        if (it.segs_per_row == 1) {
            row0 = k * it.rows_per_tile;
            nrows = it.h - row0 < it.rows_per_tile ? it.h - row0 : it.rows_per_tile;
            x0 = 0;
            npix = it.w;
        } else {
            row0 = k / it.segs_per_row;
            nrows = 1;
            x0 = (k - row0 * it.segs_per_row) * BMP_TILE_PIXELS;
            npix = it.w - x0 < BMP_TILE_PIXELS ? it.w - x0 : BMP_TILE_PIXELS;
        }
        const uint8_t *src = b.in_base + it.src;
        uint8_t *dst = b.out_base + it.dst;
        if (npix >= 1024 || nrows == 1) {
            // wide rows: the whole CTA sweeps one row (segment) at a time
            for (uint32_t r = 0; r < nrows; r++) {
                const uint32_t sr = row0 + r, dr = it.flip ? it.h - 1 - sr : sr;
                bmp_run(src + ((uint64_t)sr * it.w + x0) * 4, dst + ((uint64_t)dr * it.w + x0) * 4, npix, threadIdx.x, BMP_THREADS);
            }
        } else {
            // narrow rows: one warp per row
            for (uint32_t r = threadIdx.x >> 5; r < nrows; r += BMP_THREADS / 32) {
                const uint32_t sr = row0 + r, dr = it.flip ? it.h - 1 - sr : sr;
                bmp_run(src + ((uint64_t)sr * it.w + x0) * 4, dst + ((uint64_t)dr * it.w + x0) * 4, npix, threadIdx.x & 31, 32);
            }
        }
    }
}

}  // namespace dbg
