// sprite_kernels.cuh -- sprite-sheet tiling of decoded RGBA8 images (second half).
//
// The reference's concat_pngs.c:81-100 decodes a few PNGs and hands them to concatenate_images(to_concat, n,
// &sprite_rows, &sprite_columns), a function the reference tree does not define: there is no output to be bit-exact
// with, only an intent -- n decoded images of one size become the cells of one sheet, and the caller learns the grid.
// This is that step for a batch already in HBM (typically straight out of dbg_decode_png_batch_device): image i goes
// to cell (i / columns, i % columns) of a row-major grid, cells without an image are transparent black. Plain data
// movement: every sheet pixel is written once, every image pixel read once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dbg {

struct SpriteBatch {
    const uint8_t *rgba_base;  // image i at rgba_base + rgba_off[i], w * h * 4 bytes, rows top to bottom
    const uint64_t *rgba_off;
    uint8_t *sheet;            // (columns * w) x (rows * h) RGBA8
    uint32_t n, w, h, columns, rows;
};

// One thread per group of four horizontally adjacent sheet pixels (16 bytes) when the tile width is a multiple of four
// and source and sheet are 16-byte aligned there, else per pixel. Grid-stride; consecutive threads write consecutive
// addresses of a sheet row and read consecutive addresses of an image row.
template <int PX>
__global__ void __launch_bounds__(256) sprite_tile_kernel(SpriteBatch b)
{
    const uint64_t sheet_w = (uint64_t)b.columns * b.w, units_per_row = sheet_w / PX;
    const uint64_t total = units_per_row * ((uint64_t)b.rows * b.h);
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t y = u / units_per_row, x = (u - y * units_per_row) * PX;
        const uint32_t cell_r = (uint32_t)(y / b.h), cell_c = (uint32_t)(x / b.w);
        const uint64_t i = (uint64_t)cell_r * b.columns + cell_c;
        uint8_t *dst = b.sheet + (y * sheet_w + x) * 4;
        if (PX == 4) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (i < b.n) {
                const uint8_t *src = b.rgba_base + b.rgba_off[i] + (((y - (uint64_t)cell_r * b.h) * b.w) + (x - (uint64_t)cell_c * b.w)) * 4;
                v = *reinterpret_cast<const uint4 *>(src);
            }
            *reinterpret_cast<uint4 *>(dst) = v;
        } else {
            uint32_t v = 0;
            if (i < b.n) {
                const uint8_t *src = b.rgba_base + b.rgba_off[i] + (((y - (uint64_t)cell_r * b.h) * b.w) + (x - (uint64_t)cell_c * b.w)) * 4;
                v = *reinterpret_cast<const uint32_t *>(src);
            }
            *reinterpret_cast<uint32_t *>(dst) = v;
        }
    }
}

// aligned16: every rgba_base + rgba_off[i] and the sheet are 16-byte aligned (the caller checked)
static inline int sprite_launch(const SpriteBatch &b, bool aligned16, int sm_count, cudaStream_t s)
{
    const uint64_t px = (uint64_t)b.columns * b.w * ((uint64_t)b.rows * b.h);
    if (px == 0) return 0;
    const bool wide = aligned16 && b.w % 4 == 0;
    const uint64_t units = wide ? px / 4 : px;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((units + 255) / 256, (uint64_t)sm_count * 16);
    if (wide) sprite_tile_kernel<4><<<grid, 256, 0, s>>>(b);
    else sprite_tile_kernel<1><<<grid, 256, 0, s>>>(b);
    return (int)cudaGetLastError();
}

}  // namespace dbg
