"""ctypes binding of oracle/liboracle.so = the plain-C restatement
(oracle/debig_oracle.c). TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH) or os.path.exists(os.path.join(_HERE, "debig_oracle.c"))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "debig_oracle.c")
    if not os.path.exists(_PATH) or os.path.getmtime(src) > os.path.getmtime(_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    L = C.CDLL(_PATH, mode=C.RTLD_LOCAL)
    L.oracle_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.oracle_inflate.restype = None
    L.oracle_decode_gz.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint32)]
    L.oracle_decode_gz.restype = None
    L.oracle_png_dims.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint8)]
    L.oracle_png_dims.restype = None
    L.oracle_decode_png.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint8)]
    L.oracle_decode_png.restype = None
    L.oracle_bmp_dims.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint8)]
    L.oracle_bmp_dims.restype = None
    L.oracle_decode_bmp.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, C.POINTER(C.c_uint8)]
    L.oracle_decode_bmp.restype = None
    L.oracle_encode_bmp.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.POINTER(C.c_uint32), C.c_int64]
    L.oracle_encode_bmp.restype = None
    _lib = L
    return L


def inflate(data: bytes, cap: int):
    L = lib()
    ib = C.create_string_buffer(bytes(data) + b"\0" * 16, len(data) + 16)
    ob = C.create_string_buffer(cap + 16)
    n, g = C.c_uint64(0), C.c_uint32(0)
    L.oracle_inflate(ib, len(data), len(data) + 16, ob, cap, C.byref(n), C.byref(g))
    return int(g.value), ob.raw[: n.value] if g.value else b""


def decode_gz(data: bytes, cap: int):
    L = lib()
    ib = C.create_string_buffer(bytes(data) + b"\0" * 16, len(data) + 16)
    ob = C.create_string_buffer(cap + 16)
    n, g = C.c_uint64(0), C.c_uint32(0)
    L.oracle_decode_gz(ib, len(data), ob, cap, C.byref(n), C.byref(g))
    return int(g.value), ob.raw[: n.value] if g.value else b""


def png_dims(data: bytes):
    L = lib()
    ib = C.create_string_buffer(bytes(data), max(len(data), 1))
    w, h, g = C.c_uint32(0), C.c_uint32(0), C.c_uint8(0)
    L.oracle_png_dims(ib, len(data), C.byref(w), C.byref(h), C.byref(g))
    return int(g.value), int(w.value), int(h.value)


def decode_png(data: bytes, rgb_as_reference: bool = True):
    L = lib()
    g0, w, h = png_dims(data)
    if not g0 or w * h * 4 >= 1 << 32:
        return 0, 0, 0, b""
    ib = C.create_string_buffer(bytes(data) + b"\0" * 16, len(data) + 16)
    ob = C.create_string_buffer(w * h * 4 + 16)
    g = C.c_uint8(0)
    L.oracle_decode_png(ib, len(data), ob, w * h * 4, 1 if rgb_as_reference else 0, C.byref(g))
    return int(g.value), w, h, ob.raw[: w * h * 4] if g.value else b""


def bmp_dims(data: bytes):
    L = lib()
    ib = C.create_string_buffer(bytes(data), max(len(data), 1))
    w, h, g = C.c_uint32(0), C.c_uint32(0), C.c_uint8(0)
    L.oracle_bmp_dims(ib, len(data), C.byref(w), C.byref(h), C.byref(g))
    return int(g.value), int(w.value), int(h.value)


def decode_bmp(data: bytes, out_size: int = None):
    """Returns (good, w, h, rgba). out_size defaults to w*h*4."""
    L = lib()
    g0, w, h = bmp_dims(data)
    n = w * h * 4 if out_size is None else out_size
    if n >= 1 << 32:
        return 0, w, h, b""
    ib = C.create_string_buffer(bytes(data), max(len(data), 1))
    ob = C.create_string_buffer(max(n, 1))
    g = C.c_uint8(0)
    L.oracle_decode_bmp(ib, len(data), ob, n, C.byref(g))
    return int(g.value), w, h, ob.raw[: min(n, w * h * 4)] if g.value else b""


def encode_bmp(rgba: bytes, w: int, h: int):
    """Returns (reported size, the bytes actually written = size - 1)."""
    L = lib()
    ib = C.create_string_buffer(bytes(rgba), max(len(rgba), 1))
    cap = 54 + len(rgba) + 1
    ob = C.create_string_buffer(cap)
    n = C.c_uint32(0)
    L.oracle_encode_bmp(ib, len(rgba), w, h, ob, C.byref(n), cap)
    return int(n.value), ob.raw[: max(n.value - 1, 0)]
