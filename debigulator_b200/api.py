"""ctypes binding of libdebigulator_b200.so (include/debigulator_b200.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdebigulator_b200.so")
_lib = None

STATUS_NAMES = {
    0: "ok", 1: "cap_lt_input", 2: "input_too_small", 3: "too_large", 4: "stored_len", 5: "bad_table",
    6: "bad_code", 7: "bad_symbol", 8: "bad_distance", 9: "out_overflow", 10: "truncated", 11: "bad_repeat",
    12: "container", 13: "crc", 14: "filter", 15: "short_stream", 16: "checksum",
}
KIND_INFLATE, KIND_GZ, KIND_PNG = 0, 1, 2


class DebigulatorError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


def load_library():
    """Loads the CUDA library. There is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise DebigulatorError(
            f"{_LIB_PATH} is missing: build it with `python -m debigulator_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(_LIB_PATH, mode=C.RTLD_LOCAL)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.dbg_version.restype = i32
    L.dbg_device_count.restype = i32
    L.dbg_create.argtypes = [i32]
    L.dbg_create.restype = vp
    L.dbg_destroy.argtypes = [vp]
    L.dbg_destroy.restype = None
    L.dbg_last_error.argtypes = [vp]
    L.dbg_last_error.restype = C.c_char_p
    L.dbg_ctx_device.argtypes = [vp]
    L.dbg_ctx_device.restype = i32
    L.dbg_kernel_launches.argtypes = [vp]
    L.dbg_kernel_launches.restype = u64
    L.dbg_set_verify.argtypes = [vp, i32]
    L.dbg_set_verify.restype = i32
    L.dbg_profile_enable.argtypes = [vp, i32]
    L.dbg_profile_enable.restype = i32
    L.dbg_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
    L.dbg_profile_read.restype = i32
    L.dbg_profile_read_tag.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(u64)]
    L.dbg_profile_read_tag.restype = i32
    L.dbg_trim.argtypes = [vp]
    L.dbg_trim.restype = i32
    L.dbg_multi_create.argtypes = [i32, C.POINTER(i32)]
    L.dbg_multi_create.restype = vp
    L.dbg_multi_destroy.argtypes = [vp]
    L.dbg_multi_destroy.restype = None
    L.dbg_multi_device_count.argtypes = [vp]
    L.dbg_multi_device_count.restype = i32
    L.dbg_multi_ctx.argtypes = [vp, i32]
    L.dbg_multi_ctx.restype = vp
    L.dbg_multi_last_error.argtypes = [vp]
    L.dbg_multi_last_error.restype = C.c_char_p
    L.dbg_decode_batch_packed_multi.argtypes = [vp, i32, u64] + [vp] * 9
    L.dbg_decode_batch_packed_multi.restype = i32
    L.dbg_multi_partition.argtypes = [i32, i32, u64] + [vp] * 7
    L.dbg_multi_partition.restype = i32
    L.dbg_tile_sprites_device.argtypes = [vp, u64, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, i32, vp, u64, C.POINTER(C.c_uint32),
                                          C.POINTER(C.c_uint32), vp]
    L.dbg_tile_sprites_device.restype = i32
    L.dbg_tile_sprites.argtypes = [vp, u64, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, u64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.dbg_tile_sprites.restype = i32
    L.dbg_pipe_create.argtypes = [i32, i32]
    L.dbg_pipe_create.restype = vp
    L.dbg_pipe_destroy.argtypes = [vp]
    L.dbg_pipe_destroy.restype = None
    L.dbg_pipe_depth.argtypes = [vp]
    L.dbg_pipe_depth.restype = i32
    L.dbg_pipe_ctx.argtypes = [vp, i32]
    L.dbg_pipe_ctx.restype = vp
    L.dbg_pipe_submit.argtypes = [vp, i32, u64] + [vp] * 8
    L.dbg_pipe_submit.restype = C.c_int64
    L.dbg_pipe_wait.argtypes = [vp, C.c_int64]
    L.dbg_pipe_wait.restype = i32
    L.dbg_bsplit_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.dbg_bsplit_stats.restype = i32
    L.dbg_fx_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.dbg_fx_stats.restype = i32
    L.dbg_lane_stats.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.dbg_lane_stats.restype = i32
    L.dbg_synchronize.argtypes = [vp]
    L.dbg_synchronize.restype = i32
    for name in ("dbg_inflate_batch", "dbg_decode_gz_batch"):
        f = getattr(L, name)
        f.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp]
        f.restype = i32
    L.dbg_decode_png_batch.argtypes = [vp, u64, vp, vp, vp, vp, vp]
    L.dbg_decode_png_batch.restype = i32
    L.dbg_decode_batch_packed.argtypes = [vp, i32, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dbg_decode_batch_packed.restype = i32
    for name in ("dbg_inflate_batch_device", "dbg_decode_gz_batch_device"):
        f = getattr(L, name)
        f.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        f.restype = i32
    L.dbg_png_scratch_bytes.argtypes = [u64, u64, u64]
    L.dbg_png_scratch_bytes.restype = u64
    L.dbg_decode_png_batch_device.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, u64, u64, vp]
    L.dbg_decode_png_batch_device.restype = i32
    L.dbg_decode_bmp_batch.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp]
    L.dbg_decode_bmp_batch.restype = i32
    L.dbg_encode_bmp_batch.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dbg_encode_bmp_batch.restype = i32
    L.dbg_decode_bmp_batch_device.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dbg_decode_bmp_batch_device.restype = i32
    L.dbg_encode_bmp_batch_device.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dbg_encode_bmp_batch_device.restype = i32
    L.get_BMP_width_height.argtypes = [vp, u64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    L.get_BMP_width_height.restype = None
    L.decode_png_get_width_height.argtypes = [vp, u64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                              C.POINTER(C.c_uint8)]
    L.decode_png_get_width_height.restype = None
    _lib = L
    return L


def png_get_width_height(data: bytes):
    """decode_png_get_width_height (decode_png.h:69-75): returns (good, w, h). Host only."""
    L = load_library()
    w, h, g = C.c_uint32(0), C.c_uint32(0), C.c_uint8(0)
    buf = C.create_string_buffer(bytes(data), len(data))
    L.decode_png_get_width_height(buf, len(data), C.byref(w), C.byref(h), C.byref(g))
    return int(g.value), int(w.value), int(h.value)


def bmp_get_width_height(data: bytes):
    """get_BMP_width_height (decode_bmp.h:14-19): returns (good, w, h). Host only."""
    L = load_library()
    w, h, g = C.c_uint32(0), C.c_uint32(0), C.c_uint8(0)
    buf = C.create_string_buffer(bytes(data), max(len(data), 1))
    L.get_BMP_width_height(buf, len(data), C.byref(w), C.byref(h), C.byref(g))
    return int(g.value), int(w.value), int(h.value)


def _ptr(t):
    """Device / host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


class Context:
    """One decode context per GPU (dbg_create). Raises if no CUDA device is usable."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = self.L.dbg_create(int(device))
        if not self.h:
            raise DebigulatorError("dbg_create failed: " + self.L.dbg_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.dbg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise DebigulatorError(f"{what} failed ({rc}): " + self.L.dbg_last_error(self.h).decode())

    @property
    def kernel_launches(self):
        return int(self.L.dbg_kernel_launches(self.h))

    def set_verify(self, on=True):
        """Opt-in gzip CRC32 / ISIZE verification (dbg_set_verify)."""
        self._check(self.L.dbg_set_verify(self.h, 1 if on else 0), "dbg_set_verify")

    def profile_enable(self, on=True):
        self._check(self.L.dbg_profile_enable(self.h, 1 if on else 0), "dbg_profile_enable")

    def profile_read(self):
        """(total inflate-kernel ms, launches) since profile_enable."""
        ms, n = C.c_double(0), C.c_uint64(0)
        self._check(self.L.dbg_profile_read(self.h, C.byref(ms), C.byref(n)), "dbg_profile_read")
        return float(ms.value), int(n.value)

    PROF_INFLATE, PROF_FX_SIZES, PROF_FX_EXPAND, PROF_PNG_SCAN, PROF_PNG_UNFILTER = range(5)

    def profile_read_tag(self, tag):
        """(total ms, brackets) of one kernel group (PROF_*) since profile_enable."""
        ms, n = C.c_double(0), C.c_uint64(0)
        self._check(self.L.dbg_profile_read_tag(self.h, tag, C.byref(ms), C.byref(n)), "dbg_profile_read_tag")
        return float(ms.value), int(n.value)

    def tile_sprites(self, images, w, h, columns=0):
        """Sprite sheet of n decoded RGBA8 images (bytes, w * h * 4 each): (sheet bytes, rows, columns) (dbg_tile_sprites)."""
        n = len(images)
        cols = columns or next(c for c in range(1, n + 2) if c * c >= n)
        cols = min(cols, n)
        rows = (n + cols - 1) // cols
        bufs = [C.create_string_buffer(bytes(im), len(im)) for im in images]
        ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
        sheet = C.create_string_buffer(cols * w * rows * h * 4)
        r, c = C.c_uint32(0), C.c_uint32(0)
        self._check(self.L.dbg_tile_sprites(self.h, n, ptrs, w, h, columns, sheet, len(sheet.raw), C.byref(r), C.byref(c)), "dbg_tile_sprites")
        return sheet.raw, int(r.value), int(c.value)

    def tile_sprites_device(self, d_rgba, d_off, n, w, h, columns, aligned16, d_sheet, stream=None):
        """Device form (torch CUDA tensors; offsets as int64): returns (rows, columns)."""
        r, c = C.c_uint32(0), C.c_uint32(0)
        self._check(self.L.dbg_tile_sprites_device(self.h, n, _ptr(d_rgba), _ptr(d_off), w, h, columns, 1 if aligned16 else 0, _ptr(d_sheet),
                                                   d_sheet.numel(), C.byref(r), C.byref(c), stream), "dbg_tile_sprites_device")
        return int(r.value), int(c.value)

    def trim(self):
        """Releases the grow-only scratch of the context (dbg_trim)."""
        self._check(self.L.dbg_trim(self.h), "dbg_trim")

    def bsplit_stats(self):
        """(streams decoded by the block-split path, streams handed back to the warp-per-stream kernel)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.dbg_bsplit_stats(self.h, C.byref(a), C.byref(b)), "dbg_bsplit_stats")
        return int(a.value), int(b.value)

    def fx_stats(self):
        """(streams decoded by the lane-serial fixed-block path, streams handed back, extra decode runs)."""
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.dbg_fx_stats(self.h, C.byref(a), C.byref(b), C.byref(c)), "dbg_fx_stats")
        return int(a.value), int(b.value), int(c.value)

    def lane_stats(self):
        """dbg_lane_stats: block-split path (rounds, with end-of-block, without, chunks expanded from tokens, chunks decoded
        twice) + warp-per-stream kernel (rounds, with end-of-block, without)."""
        v = (C.c_uint32 * 8)()
        self._check(self.L.dbg_lane_stats(self.h, v), "dbg_lane_stats")
        return tuple(int(x) for x in v)

    def synchronize(self):
        self._check(self.L.dbg_synchronize(self.h), "dbg_synchronize")

    # ---- host-buffer batches (lists of bytes in, lists of bytes out) ----------
    def _pointer_batch(self, fn, items, caps):
        n = len(items)
        bufs = [C.create_string_buffer(bytes(d), max(len(d), 1)) for d in items]
        outs = [C.create_string_buffer(max(int(c), 1)) for c in caps]
        in_p = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
        out_p = (C.c_void_p * n)(*[C.addressof(b) for b in outs])
        in_sz = (C.c_uint64 * n)(*[len(d) for d in items])
        cap = (C.c_uint64 * n)(*[int(c) for c in caps])
        out_sz = (C.c_uint64 * n)()
        return n, bufs, outs, in_p, out_p, in_sz, cap, out_sz

    def inflate_batch(self, streams, caps):
        """dbg_inflate_batch: returns [(good, bytes)]."""
        n, bufs, outs, in_p, out_p, in_sz, cap, out_sz = self._pointer_batch(None, streams, caps)
        good = (C.c_uint32 * n)()
        self._check(self.L.dbg_inflate_batch(self.h, n, in_p, in_sz, out_p, cap, out_sz, good), "dbg_inflate_batch")
        return [(int(good[i]), outs[i].raw[: out_sz[i]] if good[i] else b"") for i in range(n)]

    def decode_gz_batch(self, members, caps):
        n, bufs, outs, in_p, out_p, in_sz, cap, out_sz = self._pointer_batch(None, members, caps)
        good = (C.c_uint32 * n)()
        self._check(self.L.dbg_decode_gz_batch(self.h, n, in_p, in_sz, out_p, cap, out_sz, good), "dbg_decode_gz_batch")
        return [(int(good[i]), outs[i].raw[: out_sz[i]] if good[i] else b"") for i in range(n)]

    def decode_png_batch(self, files):
        """dbg_decode_png_batch: returns [(good, w, h, rgba bytes)]."""
        dims = [png_get_width_height(f) for f in files]
        caps = [(w * h * 4 if g else 0) for g, w, h in dims]
        n, bufs, outs, in_p, out_p, in_sz, cap, _ = self._pointer_batch(None, files, caps)
        good = (C.c_uint8 * n)()
        self._check(self.L.dbg_decode_png_batch(self.h, n, in_p, in_sz, out_p, cap, good), "dbg_decode_png_batch")
        return [(int(good[i]), dims[i][1], dims[i][2], outs[i].raw[: caps[i]] if good[i] else b"") for i in range(n)]

    def decode_bmp_batch(self, files, caps=None):
        """dbg_decode_bmp_batch: returns [(good, w, h, rgba bytes)]. caps default to w*h*4 from the header."""
        if caps is None:
            dims = [bmp_get_width_height(f) for f in files]
            caps = [(w * h * 4 if g and w * h * 4 < 1 << 32 else 0) for g, w, h in dims]
        n, bufs, outs, in_p, out_p, in_sz, cap, _ = self._pointer_batch(None, files, caps)
        good = (C.c_uint8 * n)()
        w, h = (C.c_uint32 * n)(), (C.c_uint32 * n)()
        self._check(self.L.dbg_decode_bmp_batch(self.h, n, in_p, in_sz, out_p, cap, w, h, good), "dbg_decode_bmp_batch")
        return [(int(good[i]), int(w[i]), int(h[i]), outs[i].raw[: w[i] * h[i] * 4] if good[i] else b"") for i in range(n)]

    def encode_bmp_batch(self, images, caps=None):
        """dbg_encode_bmp_batch: images = [(rgba bytes, w, h)]; returns [(status, reported size, bytes written)]."""
        rgba = [im[0] for im in images]
        if caps is None:
            caps = [54 + len(r) + 1 for r in rgba]
        n, bufs, outs, in_p, out_p, in_sz, cap, out_sz = self._pointer_batch(None, rgba, caps)
        w = (C.c_uint32 * n)(*[int(im[1]) for im in images])
        h = (C.c_uint32 * n)(*[int(im[2]) for im in images])
        st = (C.c_uint32 * n)()
        self._check(self.L.dbg_encode_bmp_batch(self.h, n, in_p, in_sz, w, h, out_p, cap, out_sz, st), "dbg_encode_bmp_batch")
        return [(int(st[i]), int(out_sz[i]), outs[i].raw[: max(int(out_sz[i]) - 1, 0)] if st[i] == 0 else b"") for i in range(n)]

    # ---- packed host arenas (numpy uint8 arrays, ideally pinned) --------------
    def decode_packed(self, kind, h_in, in_off, in_size, h_out, out_off, out_cap):
        n = len(in_off)
        in_off = np.ascontiguousarray(in_off, dtype=np.uint64)
        in_size = np.ascontiguousarray(in_size, dtype=np.uint64)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out_cap = np.ascontiguousarray(out_cap, dtype=np.uint64)
        out_size = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.uint32)
        self._check(self.L.dbg_decode_batch_packed(self.h, int(kind), n, _ptr(h_in), _ptr(in_off), _ptr(in_size),
                                                   _ptr(h_out), _ptr(out_off), _ptr(out_cap), _ptr(out_size),
                                                   _ptr(status)), "dbg_decode_batch_packed")
        return out_size, status

    # ---- device-resident batches (torch CUDA tensors; offsets as int64) -------
    def inflate_device(self, d_in, in_off, in_size, d_out, out_off, out_cap, out_size, status, order=None,
                       stream=None, gz=False):
        fn = self.L.dbg_decode_gz_batch_device if gz else self.L.dbg_inflate_batch_device
        self._check(fn(self.h, in_off.numel(), _ptr(d_in), _ptr(in_off), _ptr(in_size), _ptr(d_out), _ptr(out_off),
                       _ptr(out_cap), _ptr(out_size), _ptr(status), _ptr(order), stream),
                    "dbg_decode_gz_batch_device" if gz else "dbg_inflate_batch_device")

    def png_device(self, d_in, in_off, in_size, d_out, out_off, out_cap, status, total_in, total_rgba, stream=None):
        self._check(self.L.dbg_decode_png_batch_device(self.h, in_off.numel(), _ptr(d_in), _ptr(in_off), _ptr(in_size),
                                                       _ptr(d_out), _ptr(out_off), _ptr(out_cap), _ptr(status),
                                                       int(total_in), int(total_rgba), stream),
                    "dbg_decode_png_batch_device")

    def bmp_decode_device(self, d_in, in_off, in_size, d_out, out_off, out_cap, out_size, status, width=None,
                          height=None, stream=None):
        self._check(self.L.dbg_decode_bmp_batch_device(self.h, in_off.numel(), _ptr(d_in), _ptr(in_off), _ptr(in_size),
                                                       _ptr(d_out), _ptr(out_off), _ptr(out_cap), _ptr(out_size),
                                                       _ptr(width), _ptr(height), _ptr(status), stream),
                    "dbg_decode_bmp_batch_device")

    def bmp_encode_device(self, d_rgba, rgba_off, rgba_size, width, height, d_out, out_off, out_cap, out_size, status,
                          stream=None):
        self._check(self.L.dbg_encode_bmp_batch_device(self.h, rgba_off.numel(), _ptr(d_rgba), _ptr(rgba_off),
                                                       _ptr(rgba_size), _ptr(width), _ptr(height), _ptr(d_out),
                                                       _ptr(out_off), _ptr(out_cap), _ptr(out_size), _ptr(status), stream),
                    "dbg_encode_bmp_batch_device")


def partition(n_devices, kind, in_off, in_size, out_off, out_cap, h_in=None):
    """dbg_multi_partition: (device index per item, estimated cost per device, runs). Pure host code, needs no GPU."""
    L = load_library()
    n = len(in_off)
    a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (in_off, in_size, out_off, out_cap)]
    dev = np.zeros(n, dtype=np.uint32)
    cost = np.zeros(n_devices, dtype=np.uint64)
    runs = L.dbg_multi_partition(int(n_devices), int(kind), n, _ptr(h_in), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]),
                                 _ptr(dev), _ptr(cost))
    if runs < 0:
        raise DebigulatorError("dbg_multi_partition: bad arguments")
    return dev, cost, runs


class MultiContext:
    """Several GPUs of one box behind one call (dbg_multi_*): one context, stream set and host thread per device."""

    def __init__(self, n_devices=0, device_ids=None):
        self.L = load_library()
        ids = (C.c_int * len(device_ids))(*device_ids) if device_ids else None
        self.h = self.L.dbg_multi_create(int(n_devices), ids)
        if not self.h:
            raise DebigulatorError("dbg_multi_create failed: " + (self.L.dbg_last_error(None) or b"").decode())
        self.n_devices = self.L.dbg_multi_device_count(self.h)

    def close(self):
        if self.h:
            self.L.dbg_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fx_stats(self, k):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.L.dbg_fx_stats(self.L.dbg_multi_ctx(self.h, k), C.byref(a), C.byref(b), C.byref(c))
        return int(a.value), int(b.value), int(c.value)

    def decode_packed(self, kind, h_in, in_off, in_size, h_out, out_off, out_cap):
        """Returns (out_size, status, device_of_item)."""
        n = len(in_off)
        a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (in_off, in_size, out_off, out_cap)]
        out_size = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.uint32)
        dev = np.zeros(n, dtype=np.uint32)
        rc = self.L.dbg_decode_batch_packed_multi(self.h, int(kind), n, _ptr(h_in), _ptr(a[0]), _ptr(a[1]), _ptr(h_out), _ptr(a[2]),
                                                  _ptr(a[3]), _ptr(out_size), _ptr(status), _ptr(dev))
        if rc != 0:
            raise DebigulatorError("dbg_decode_batch_packed_multi failed (%d): %s" % (rc, (self.L.dbg_multi_last_error(self.h) or b"").decode()))
        return out_size, status, dev


class Pipe:
    """Several packed batches in flight on one GPU (dbg_pipe_*): submit() returns a ticket at once, wait() the batch's
    (out_size, status). The arrays handed to submit() must stay alive and untouched until wait() has returned."""

    def __init__(self, device=0, depth=2):
        self.L = load_library()
        self.h = self.L.dbg_pipe_create(int(device), int(depth))
        if not self.h:
            raise DebigulatorError("dbg_pipe_create failed: " + (self.L.dbg_last_error(None) or b"").decode())
        self.depth = self.L.dbg_pipe_depth(self.h)
        self._jobs = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.dbg_pipe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def kernel_launches(self):
        return sum(int(self.L.dbg_kernel_launches(self.L.dbg_pipe_ctx(self.h, k))) for k in range(self.depth))

    def submit(self, kind, h_in, in_off, in_size, h_out, out_off, out_cap):
        n = len(in_off)
        a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (in_off, in_size, out_off, out_cap)]
        out_size = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.uint32)
        t = int(self.L.dbg_pipe_submit(self.h, int(kind), n, _ptr(h_in), _ptr(a[0]), _ptr(a[1]), _ptr(h_out), _ptr(a[2]), _ptr(a[3]),
                                       _ptr(out_size), _ptr(status)))
        if t < 0:
            raise DebigulatorError("dbg_pipe_submit failed (%d)" % t)
        self._jobs[t] = (a, out_size, status, h_in, h_out)
        return t

    def wait(self, ticket):
        rc = self.L.dbg_pipe_wait(self.h, int(ticket))
        a, out_size, status, _, _ = self._jobs.pop(ticket)
        if rc != 0:
            raise DebigulatorError("packed batch of ticket %d failed (%d)" % (ticket, rc))
        return out_size, status
