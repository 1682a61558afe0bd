"""ctypes binding of oracle/_ref/libref.so = the UNMODIFIED reference C sources
(inflate.c, decode_png.c) compiled by oracle/Makefile, plus oracle/ref_shim.c.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref.so")
_lib = None

PAD_IN = 64      # readable bytes after the input (inflate.c:252-256 over-read)
SLACK_OUT = 2048  # writable bytes after the output capacity (Q4 4x over-copy)


def available() -> bool:
    return os.path.exists(_PATH) or os.path.isdir("/root/reference/src")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_PATH) and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    L = C.CDLL(_PATH, mode=C.RTLD_LOCAL)
    u8p = C.POINTER(C.c_uint8)
    L.ref_init.restype = None
    L.ref_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                              C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.ref_inflate.restype = None
    L.ref_decode_gz.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.ref_decode_gz.restype = None
    L.ref_png_dims.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32),
                               C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)]
    L.ref_png_dims.restype = None
    L.ref_decode_png.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                 C.POINTER(C.c_uint8)]
    L.ref_decode_png.restype = None
    L.ref_stb_png.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.POINTER(C.c_int)]
    L.ref_stb_png.restype = C.c_void_p
    L.ref_stb_zlib.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.ref_stb_zlib.restype = C.c_void_p
    L.ref_free.argtypes = [C.c_void_p]
    L.ref_free.restype = None
    # decode_bmp.c is self-contained: its three entry points are bound directly
    L.get_BMP_width_height.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_uint8)]
    L.get_BMP_width_height.restype = None
    L.decode_BMP.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, C.POINTER(C.c_uint8)]
    L.decode_BMP.restype = None
    L.encode_BMP.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32),
                             C.c_int64]
    L.encode_BMP.restype = None
    L.ref_init()
    _lib = L
    return L


def _inbuf(data: bytes):
    return C.create_string_buffer(bytes(data) + b"\0" * PAD_IN, len(data) + PAD_IN)


def inflate(data: bytes, cap: int):
    """Reference inflate(): returns (good, bytes[0:final_size])."""
    L = lib()
    ib = _inbuf(data)
    ob = C.create_string_buffer(cap + SLACK_OUT)
    n = C.c_uint64(0)
    g = C.c_uint32(0)
    L.ref_inflate(ib, len(data), ob, cap, C.byref(n), C.byref(g))
    return int(g.value), ob.raw[: n.value] if g.value else b""


def decode_gz(data: bytes, cap: int):
    L = lib()
    ib = _inbuf(data)
    ob = C.create_string_buffer(cap + SLACK_OUT)
    n = C.c_uint64(0)
    g = C.c_uint32(0)
    L.ref_decode_gz(ib, len(data), ob, cap, C.byref(n), C.byref(g))
    return int(g.value), ob.raw[: n.value] if g.value else b""


def png_dims(data: bytes):
    L = lib()
    ib = _inbuf(data)
    w = C.c_uint32(0)
    h = C.c_uint32(0)
    g = C.c_uint8(0)
    L.ref_png_dims(ib, len(data), C.byref(w), C.byref(h), C.byref(g))
    return int(g.value), int(w.value), int(h.value)


def decode_png(data: bytes):
    """Reference decode_png(): returns (good, w, h, rgba bytes)."""
    L = lib()
    g0, w, h = png_dims(data)
    if not g0:
        return 0, 0, 0, b""
    ib = _inbuf(data)
    ob = C.create_string_buffer(w * h * 4 + 16)
    g = C.c_uint8(0)
    L.ref_decode_png(ib, len(data), ob, w * h * 4, C.byref(g))
    return int(g.value), w, h, ob.raw[: w * h * 4] if g.value else b""


def stb_png(pixels: bytes, w: int, h: int, comp: int = 4, filt: int = -1) -> bytes:
    """PNG bytes from the reference's vendored stb_write.h (stb_write.h:1128)."""
    L = lib()
    n = C.c_int(0)
    pb = C.create_string_buffer(bytes(pixels), len(pixels))
    p = L.ref_stb_png(pb, w, h, comp, filt, C.byref(n))
    out = C.string_at(p, n.value)
    L.ref_free(p)
    return out


def stb_zlib(data: bytes, quality: int = 8) -> bytes:
    L = lib()
    n = C.c_int(0)
    pb = C.create_string_buffer(bytes(data), len(data))
    p = L.ref_stb_zlib(pb, len(data), quality, C.byref(n))
    out = C.string_at(p, n.value)
    L.ref_free(p)
    return out


def bmp_dims(data: bytes):
    """Reference get_BMP_width_height (decode_bmp.c:53). Needs >= 54 bytes (the reference asserts)."""
    L = lib()
    if len(data) < 54:
        return 0, 0, 0
    ib = _inbuf(data)
    w, h, g = C.c_uint32(0), C.c_uint32(0), C.c_uint8(0)
    L.get_BMP_width_height(ib, len(data), C.byref(w), C.byref(h), C.byref(g))
    return int(g.value), int(w.value), int(h.value)


def _bmp_in_bounds(data: bytes, out_size: int) -> bool:
    """True when the reference's pixel loop (decode_bmp.c:260-291) stays inside both buffers."""
    import struct
    if len(data) < 54:
        return False
    off = struct.unpack_from("<I", data, 10)[0]
    w, h = struct.unpack_from("<ii", data, 18)
    if w < 0 or h == -(1 << 31):
        return False
    need = w * abs(h) * 4
    return need < (1 << 32) and need <= out_size and off + need <= len(data)


def decode_bmp(data: bytes, out_size: int = None):
    """Reference decode_BMP (decode_bmp.c:105): returns (good, w, h, rgba). Inputs on which the reference
    would index out of bounds are reported as good = 0 without calling it."""
    L = lib()
    g0, w, h = bmp_dims(data)
    n = w * h * 4 if out_size is None else out_size
    if len(data) < 54 or n >= 1 << 32:
        return 0, w, h, b""
    ib = _inbuf(data)
    # the header checks run before any pixel is touched, so a rejected file is safe to pass through
    import struct
    hdr_ok = (data[:2] == b"BM" and struct.unpack_from("<I", data, 14)[0] in (40, 108)
              and struct.unpack_from("<HH", data, 26) == (1, 32)
              and not (((struct.unpack_from("<I", data, 10)[0] + struct.unpack_from("<I", data, 2)[0]) & 0xffffffff) + 14 < len(data)))
    if hdr_ok and not _bmp_in_bounds(data, n):
        return 0, w, h, b""
    ob = C.create_string_buffer(max(n, 1) + 16)
    g = C.c_uint8(0)
    L.decode_BMP(ib, len(data), ob, n, C.byref(g))
    return int(g.value), w, h, ob.raw[: min(n, w * h * 4)] if g.value else b""


def encode_bmp(rgba: bytes, w: int, h: int):
    """Reference encode_BMP (decode_bmp.c:297): returns (reported size, bytes actually written = size - 1)."""
    L = lib()
    assert len(rgba) % 4 == 0
    ib = _inbuf(rgba)
    cap = 54 + len(rgba) + 1
    ob = C.create_string_buffer(cap + 16)
    n = C.c_uint32(0)
    L.encode_BMP(ib, len(rgba), w, h, ob, C.byref(n), cap)
    return int(n.value), ob.raw[: max(n.value - 1, 0)]
