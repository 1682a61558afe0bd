/*
 * debigulator_b200.h -- batched C-ABI of the B200-native inflate / gzip / PNG
 * decode path (libdebigulator_b200.so).
 *
 * The scalar, reference-compatible entry points live beside this file in
 * inflate.h, decode_png.h and decode_gz.h (same names and signatures as the
 * reference's src/inflate.h:22-60, src/decode_png.h:43-103 and
 * src/decode_gz.h:23-38). This header adds what the reference does not have:
 * batches of independent streams / images decoded by hand-written sm_100a CUDA
 * kernels. There is NO CPU fallback: without a usable CUDA device dbg_create()
 * returns NULL and every entry point reports DBG_ERR_NO_DEVICE.
 *
 * Conventions shared with the reference:
 *   - success of an item is an out-param flag, 1 = good, 0 = failed
 *     (uint32_t for inflate / gzip as in inflate.h:59, uint8_t for PNG as in
 *     decode_png.h:103); the int return value of a batch call only reports
 *     infrastructure failures (CUDA errors, bad arguments).
 *   - inflate items follow inflate()'s contract: raw DEFLATE in, capacity
 *     out_cap[i] must be >= in_size[i] and in_size[i] >= 5 (inflate.c:826-844),
 *     out_size[i] is the reference's *final_recipient_size.
 *   - gzip items follow decode_gz()'s framing (decode_gz.c:131-233, 270): 10
 *     byte header, optional FNAME, deflate payload = rest - 8; CRC32 / ISIZE are
 *     not verified (decode_gz.c:281-297).
 *   - PNG items follow decode_png(): output is RGBA8, w*h*4 bytes, and
 *     rgba_size[i] must equal w*h*4 (decode_png.c:970). Unlike the reference
 *     the input buffers are NOT modified.
 *
 * Plain C, no CUDA or torch types in any signature: device pointers and
 * streams cross this boundary as void* / uint8_t*.
 */
#ifndef DEBIGULATOR_B200_H
#define DEBIGULATOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dbg_ctx dbg_ctx;

enum {
    DBG_OK = 0,
    DBG_ERR_NO_DEVICE = -1, /* no CUDA device / driver: there is no CPU path */
    DBG_ERR_CUDA = -2,      /* a CUDA call failed; see dbg_last_error() */
    DBG_ERR_ARG = -3,       /* NULL / inconsistent arguments */
    DBG_ERR_NOMEM = -4      /* host or device allocation failed */
};

/* Per-item status codes written by the *_device entry points (0 = good). The
 * host entry points collapse them to the reference's 1/0 `good` flag. */
enum {
    DBG_ST_OK = 0,
    DBG_ST_CAP_LT_INPUT = 1,
    DBG_ST_INPUT_TOO_SMALL = 2,
    DBG_ST_TOO_LARGE = 3,
    DBG_ST_STORED_LEN = 4,
    DBG_ST_BAD_TABLE = 5,
    DBG_ST_BAD_CODE = 6,
    DBG_ST_BAD_SYMBOL = 7,
    DBG_ST_BAD_DISTANCE = 8,
    DBG_ST_OUT_OVERFLOW = 9,
    DBG_ST_TRUNCATED = 10,
    DBG_ST_BAD_REPEAT = 11,
    DBG_ST_CONTAINER = 12,  /* gzip / PNG framing rejected the item */
    DBG_ST_CRC = 13,        /* PNG chunk CRC mismatch (decode_png.c:1341-1348) */
    DBG_ST_FILTER = 14,     /* first filter byte > 4 (decode_png.c:847-858) */
    DBG_ST_SHORT_STREAM = 15, /* inflated PNG stream shorter than h*(w*bpp+1) */
    DBG_ST_CHECKSUM = 16     /* only with dbg_set_verify(): gzip CRC32 / ISIZE or zlib Adler-32 mismatch */
};

int dbg_version(void);
int dbg_device_count(void); /* >= 0, or a DBG_ERR_* code */

/* One context per GPU (one process per GPU in multi-GPU runs, or dbg_multi below). A context owns its device scratch,
 * pinned staging buffers and streams; it is not thread-safe -- use one context per host thread. Device-resident
 * calls on different streams are ordered by the library itself (each waits, on the device, for the previous one to
 * be done with the shared scratch). */
dbg_ctx *dbg_create(int device);
void dbg_destroy(dbg_ctx *ctx);
const char *dbg_last_error(const dbg_ctx *ctx); /* ctx may be NULL */
int dbg_ctx_device(const dbg_ctx *ctx);
uint64_t dbg_kernel_launches(const dbg_ctx *ctx); /* kernels launched by this ctx so far */

/* ---- host-buffer batches: H2D + kernels + D2H inside the call ------------ */

/* Batched inflate(): replaces n calls of inflate() (inflate.h:51-60). */
int dbg_inflate_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                      uint8_t *const *out, const uint64_t *out_cap, uint64_t *out_size, uint32_t *good);

/* Batched decode_gz(): replaces n calls of decode_gz() (decode_gz.h:36-38);
 * the caller supplies the output buffers (e.g. sized from ISIZE). */
int dbg_decode_gz_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                        uint8_t *const *out, const uint64_t *out_cap, uint64_t *out_size, uint32_t *good);

/* Batched decode_png(): replaces n calls of decode_png() (decode_png.h:96-103). */
int dbg_decode_png_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                         uint8_t *const *out_rgba, const uint64_t *rgba_size, uint8_t *good);

/* Packed variants: one host input arena and one host output arena (ideally
 * pinned), items addressed by offsets. kind: 0 = raw deflate, 1 = gzip member,
 * 2 = PNG (out_cap = w*h*4; out_size[i] is set to out_cap[i] for good images).
 * A single H2D and a single D2H per wave; this is the end-to-end fast path. */
int dbg_decode_batch_packed(dbg_ctx *ctx, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                            const uint64_t *in_size, uint8_t *h_out, const uint64_t *out_off,
                            const uint64_t *out_cap, uint64_t *out_size, uint32_t *status);

/* ---- device-resident batches: everything already in HBM ------------------ */
/* All d_* pointers are device pointers on ctx's device. Input item i occupies
 * d_in[in_off[i] .. in_off[i]+in_size[i]); the bytes from (address & ~15) up to
 * the next 16-byte boundary after its end must be readable (pad the arena by
 * 16). Output item i is written at d_out + out_off[i], at most out_cap[i]
 * bytes. d_order (may be NULL) is a permutation giving the scheduling order
 * (heaviest first balances best; with NULL the library sorts the queue itself).
 * Work is enqueued on `stream` (a cudaStream_t, NULL = the context's own
 * stream) and the call returns with it still running. It may wait for `stream`
 * on the way: the chunk-parallel paths for long streams
 * read a few counters back before they size their scratch memory. */
int dbg_inflate_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                             const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                             const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_status,
                             const uint32_t *d_order, void *stream);

int dbg_decode_gz_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                               const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                               const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_status,
                               const uint32_t *d_order, void *stream);

/* PNG: d_out_cap[i] = w*h*4 as the caller computed it from
 * decode_png_get_width_height(). scratch_bytes (host value) must be at least
 * dbg_png_scratch_bytes(n, total_in_bytes, total_rgba_bytes); the context
 * allocates and keeps that scratch (compacted IDAT streams + filtered
 * scanlines). */
uint64_t dbg_png_scratch_bytes(uint64_t n, uint64_t total_in_bytes, uint64_t total_rgba_bytes);
int dbg_decode_png_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                                const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                                const uint64_t *d_out_cap, uint32_t *d_status, uint64_t total_in_bytes,
                                uint64_t total_rgba_bytes, void *stream);

/* Opt-in strictness the reference does not have: when on, gzip batches also compare CRC-32 and ISIZE of every
 * decoded member with its trailer (the reference reads the trailer but never checks it, decode_gz.c:281-297) and
 * PNG batches compare the Adler-32 of the inflated scanlines with the zlib trailer (the reference drops those
 * four bytes unread, decode_png.c:816); a mismatch reports DBG_ST_CHECKSUM / good = 0. Computed on the device.
 * Off by default, so that `good` matches the reference. */
int dbg_set_verify(dbg_ctx *ctx, int on);

/* How many streams so far were decoded chunk-parallel by the block-split path (long multi-block streams,
 * see DESIGN.md 4.3), and how many of those were handed back to the warp-per-stream kernel because a
 * hinted block boundary turned out not to be one. Diagnostics only; either pointer may be NULL. */
int dbg_bsplit_stats(const dbg_ctx *ctx, uint64_t *streams, uint64_t *fallbacks);

/* The same for the lane-serial path that decodes single fixed-Huffman-block streams (what stb_image_write emits
 * for every PNG): streams decoded there, streams handed back to the warp-per-stream kernel, and
 * chunk entry points that needed more than one decode run (strictly periodic symbol streams). Diagnostics only. */
int dbg_fx_stats(const dbg_ctx *ctx, uint64_t *streams, uint64_t *handed_back, uint64_t *extra_runs);

/* Diagnostics of the lane-parallel block decode, counted on the device (waits for it). Block-split
 * path: v[0] rounds, v[1] rounds that met end-of-block, v[2] rounds that stopped before it, v[3] chunks whose recorded
 * tokens were expanded, v[4] chunks that were Huffman-decoded a second time instead. Warp-per-stream kernel: v[5]
 * rounds, v[6] rounds that met end-of-block, v[7] rounds that did not. */
int dbg_lane_stats(const dbg_ctx *ctx, uint32_t v[8]);

/* Optional timing of the kernel groups, for roofline reports: after dbg_profile_enable(ctx, 1) every group below is
 * bracketed by CUDA events on the stream it runs on. dbg_profile_read_tag() waits for the brackets of one group and
 * returns their summed device time and their number; dbg_profile_read() does that for DBG_PROF_INFLATE and resets
 * the recording (dbg_profile_enable(ctx, 1) resets it too). */
enum {
    DBG_PROF_INFLATE = 0,      /* inflate_batch_kernel: one warp per stream */
    DBG_PROF_FX_SIZES = 1,     /* lane-serial path: head + sizes + chain kernels */
    DBG_PROF_FX_EXPAND = 2,    /* lane-serial path: tokens + expansion + resolve kernels */
    DBG_PROF_PNG_SCAN = 3,     /* PNG chunk walk, CRC-32, IDAT gather */
    DBG_PROF_PNG_UNFILTER = 4  /* PNG scanline reconstruction */
};
int dbg_profile_enable(dbg_ctx *ctx, int on);
int dbg_profile_read(dbg_ctx *ctx, double *total_ms, uint64_t *launches);
int dbg_profile_read_tag(dbg_ctx *ctx, int tag, double *total_ms, uint64_t *launches);

/* Releases the scratch memory the context keeps between calls: 16-bit cells (2 bytes per output byte of the streams
 * on the chunk-parallel paths), tokens (4 bytes per symbol; block-split path: up to 16 bytes per compressed byte,
 * capped at 24 GiB), PNG scanline buffers (1.25 x the RGBA bytes of a batch) and the staging arenas of the host API.
 * All of it is grow-only otherwise. Waits for the device. */
int dbg_trim(dbg_ctx *ctx);

/* One gzip member, decoded into memory obtained from `alloc` once the size is known (the drop-in decode_gz() of
 * decode_gz.h is built on this). cap = largest output accepted. *good = 0 leaves *out NULL. */
int dbg_decode_gz_alloc(dbg_ctx *ctx, const uint8_t *in, uint64_t in_size, uint64_t cap, void *(*alloc)(size_t),
                        uint8_t **out, uint64_t *out_size, uint32_t *good);

/* ---- several GPUs of one box, one batch ---------------------------------------------------------
 * The items of a batch are independent: the batch is cut into runs of consecutive items, the runs are dealt to the
 * devices longest-processing-time first by estimated decode time, and every device decodes its runs with its own
 * context, stream set and host thread. No exchange step, no collective. The reference's only provision for
 * concurrency is the thread_id slot (inflate.c:22-23, decode_png.c:559-560). */
typedef struct dbg_multi dbg_multi;
dbg_multi *dbg_multi_create(int n_devices, const int *device_ids); /* n_devices <= 0: all; device_ids NULL: 0..n-1 */
void dbg_multi_destroy(dbg_multi *m);
int dbg_multi_device_count(const dbg_multi *m);
dbg_ctx *dbg_multi_ctx(dbg_multi *m, int k); /* the k-th device's context (counters, tunables) */
const char *dbg_multi_last_error(const dbg_multi *m);
/* Same contract as dbg_decode_batch_packed(); device_of_item (may be NULL) receives the device index of every item. */
int dbg_decode_batch_packed_multi(dbg_multi *m, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                                  const uint64_t *in_size, uint8_t *h_out, const uint64_t *out_off, const uint64_t *out_cap,
                                  uint64_t *out_size, uint32_t *status, uint32_t *device_of_item);
/* The partition alone (pure host code; needs no GPU): device index per item, estimated cost per device (may be
 * NULL). h_in may be NULL (costs then come from the sizes only). Returns the number of runs, or DBG_ERR_ARG. */
int dbg_multi_partition(int n_devices, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                        const uint64_t *in_size, const uint64_t *out_off, const uint64_t *out_cap, uint32_t *device_of_item,
                        uint64_t *device_cost);

/* ---- sprite sheets (intent of concat_pngs.c:81-100) -----------------------------------------------
 * The reference's concat_pngs.c decodes a few PNGs and calls concatenate_images(to_concat, n, &sprite_rows,
 * &sprite_columns), which its tree does not define; this is that step for a decoded batch: n RGBA8 images of w x h
 * pixels each become the cells of one sheet of `columns` cells per row (0: ceil(sqrt(n))), image i in cell
 * (i / columns, i % columns), cells without an image transparent black. The sheet is (columns * w) x (rows * h) pixels,
 * rows = ceil(n / columns); sheet_cap must hold it. *out_rows / *out_columns (may be NULL) receive the grid.
 * Device form: d_rgba_off is a device array of byte offsets into d_rgba (any 4-byte alignment; 16-byte aligned images
 * and sheet with w % 4 == 0 take 16-byte accesses -- pass offsets_16_aligned = 1 when that holds). Returns DBG_ERR_ARG
 * when the sheet does not fit or a size is zero. */
int dbg_tile_sprites_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_rgba, const uint64_t *d_rgba_off, uint32_t w, uint32_t h,
                            uint32_t columns, int offsets_16_aligned, uint8_t *d_sheet, uint64_t sheet_cap, uint32_t *out_rows,
                            uint32_t *out_columns, void *stream);
/* Host form: n pointers to decoded images (w * h * 4 bytes each), sheet in host memory. */
int dbg_tile_sprites(dbg_ctx *ctx, uint64_t n, const uint8_t *const *rgba, uint32_t w, uint32_t h, uint32_t columns,
                     uint8_t *sheet, uint64_t sheet_cap, uint32_t *out_rows, uint32_t *out_columns);

/* ---- several packed batches in flight on one GPU ---------------------------------------------------------------
 * dbg_decode_batch_packed() returns when its batch is done, so back-to-back calls pay the ramp of every batch in full:
 * the download engine idles until the first wave's kernels have finished, the upload engine after the last wave's
 * upload. A pipe keeps `depth` batches in flight, each on a context, stream set and host thread of its own, so that one
 * batch's ramp runs under the previous batch's downloads (a caller that streams batches through the decoder: the
 * reference's counterpart is a loop over decode_gz() / decode_png(), decode_gz.c:123, decode_png.c:683).
 * dbg_pipe_submit() takes the arguments of dbg_decode_batch_packed(), returns a ticket >= 0 (or a negative DBG_ERR_*)
 * and blocks only while `depth` batches are already in flight; every buffer passed to it belongs to the pipe until
 * dbg_pipe_wait() has returned that ticket's result (the batch call's return code). Tickets may be waited for in any
 * order, once each. Submit and wait may be called from one thread or several. */
typedef struct dbg_pipe dbg_pipe;
dbg_pipe *dbg_pipe_create(int device, int depth); /* depth 1..4 */
void dbg_pipe_destroy(dbg_pipe *p);               /* waits for the batches in flight */
int dbg_pipe_depth(const dbg_pipe *p);
dbg_ctx *dbg_pipe_ctx(dbg_pipe *p, int k);        /* the k-th context (counters, tunables) */
int64_t dbg_pipe_submit(dbg_pipe *p, int kind, uint64_t n, const uint8_t *h_in, const uint64_t *in_off,
                        const uint64_t *in_size, uint8_t *h_out, const uint64_t *out_off, const uint64_t *out_cap,
                        uint64_t *out_size, uint32_t *status);
int dbg_pipe_wait(dbg_pipe *p, int64_t ticket);

/* ---- BMP (decode_bmp.h:14-36; decode_bmp.c:105-372) --------------------------
 * 32-bit BGRA BMP <-> RGBA8, the reference's decode_BMP / encode_BMP batched. Decode accepts what the
 * reference accepts ('BM', 40- or 108-byte DIB header, 1 plane, 32 bpp, either row order) and writes
 * w*h*4 bytes (out_size); inputs the reference would read or write out of bounds on are rejected
 * instead (DBG_ST_TRUNCATED / DBG_ST_OUT_OVERFLOW / DBG_ST_TOO_LARGE). Encode writes the 54 header
 * bytes + swizzled pixels and reports 54 + rgba_size + 1 like the reference (the last byte is never
 * written); rgba_size must be a multiple of 4 and out_cap at least the reported size.
 * width / height of dbg_decode_bmp_batch may be NULL. */
int dbg_decode_bmp_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *in, const uint64_t *in_size,
                         uint8_t *const *out_rgba, const uint64_t *rgba_cap, uint32_t *width, uint32_t *height,
                         uint8_t *good);
int dbg_encode_bmp_batch(dbg_ctx *ctx, uint64_t n, const uint8_t *const *rgba, const uint64_t *rgba_size,
                         const uint32_t *width, const uint32_t *height, uint8_t *const *out, const uint64_t *out_cap,
                         uint64_t *out_size, uint32_t *status);
/* Device-resident forms (d_width / d_height of the decode may be NULL). */
int dbg_decode_bmp_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_in, const uint64_t *d_in_off,
                                const uint64_t *d_in_size, uint8_t *d_out, const uint64_t *d_out_off,
                                const uint64_t *d_out_cap, uint64_t *d_out_size, uint32_t *d_width,
                                uint32_t *d_height, uint32_t *d_status, void *stream);
int dbg_encode_bmp_batch_device(dbg_ctx *ctx, uint64_t n, const uint8_t *d_rgba, const uint64_t *d_rgba_off,
                                const uint64_t *d_rgba_size, const uint32_t *d_width, const uint32_t *d_height,
                                uint8_t *d_out, const uint64_t *d_out_off, const uint64_t *d_out_cap,
                                uint64_t *d_out_size, uint32_t *d_status, void *stream);

/* Blocks until everything the context enqueued on its own stream is done. */
int dbg_synchronize(dbg_ctx *ctx);

#ifdef __cplusplus
}
#endif

#endif /* DEBIGULATOR_B200_H */
