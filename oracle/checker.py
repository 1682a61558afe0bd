"""The checker tests, smoke() and bench.py's CPU legs use: the unmodified
reference (oracle/_ref/libref.so, kind "reference") when it is available, else
the plain-C restatement (oracle/liboracle.so, kind "port").

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
from . import reflib

try:
    from . import portlib
except Exception:  # pragma: no cover
    portlib = None


def kind() -> str:
    return "reference" if reflib.available() else "port"


def _impl():
    if reflib.available():
        return reflib
    if portlib is None or not portlib.available():
        raise RuntimeError("no CPU checker available: build oracle/ (make -C oracle)")
    return portlib


def inflate(data: bytes, cap: int):
    return _impl().inflate(data, cap)


def decode_gz(data: bytes, cap: int):
    return _impl().decode_gz(data, cap)


def decode_png(data: bytes):
    return _impl().decode_png(data)


def decode_bmp(data: bytes, out_size: int = None):
    return _impl().decode_bmp(data, out_size)


def encode_bmp(rgba: bytes, w: int, h: int):
    return _impl().encode_bmp(rgba, w, h)
