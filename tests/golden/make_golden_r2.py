"""Round-2 additions to the golden vectors (tests/golden/manifest_r2.json), produced like manifest.json by running the
UNMODIFIED reference (oracle/_ref/libref.so) -- the negative / edge vectors SURVEY.md appendix E asks for that
manifest.json does not hold: zlib header checks (decode_png.c:1186, :1214-1220, :1262-1265), colour types 0 / 4 and bit
depth 16 (:987-1038, :1088), interlace (not checked in the silent build, :1115), gzip members with FCOMMENT / FEXTRA /
FHCRC set (decode_gz.c:195-233: the silent build skips FNAME only), and the four Paeth-heavy fs_*.png fixtures.
Deterministic (no os.urandom), so re-running it reproduces the file. Build container only:

    python tests/golden/make_golden_r2.py
"""
import base64
import hashlib
import io
import json
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reflib  # noqa: E402
from debigulator_b200 import corpus  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sha = lambda b: hashlib.sha256(b).hexdigest()
b64 = lambda b: base64.b64encode(b).decode()


def chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def png_file(w, h, depth, ct, interlace, stream, extra=()):
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ct, 0, 0, interlace))
    for tag, data in extra:
        out += chunk(tag, data)
    return out + chunk(b"IDAT", stream) + chunk(b"IEND", b"")


def zstream(rows, cmf=0x78, flg=None, fdict=False):
    """zlib stream with a chosen header: FLG's check bits are made valid unless flg is given verbatim."""
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = c.compress(rows) + c.flush()
    if flg is None:
        flg = 0x20 if fdict else 0
        flg += (31 - ((cmf << 8) | flg) % 31) % 31
    return bytes([cmf, flg]) + body + struct.pack(">I", zlib.adler32(rows) & 0xFFFFFFFF)


def png_vectors():
    out = []
    w = h = 12
    img = corpus.gradient_noise_rgba(w, h, 77)
    rows4 = corpus.png_filter_rows(img, -1)

    def add(name, data, note):
        good, rw, rh, o = reflib.decode_png(data)
        out.append(dict(name=name, in_b64=b64(data), good=good, w=rw, h=rh, out_sha256=sha(o), note=note))

    add("zlib_ok_control", png_file(w, h, 8, 6, 0, zstream(rows4)), "control: the same image with a regular header")
    add("zlib_cm_not_8", png_file(w, h, 8, 6, 0, zstream(rows4, cmf=0x77)), "CM = 7 with valid check bits (decode_png.c:1186)")
    add("zlib_bad_fcheck", png_file(w, h, 8, 6, 0, zstream(rows4, flg=0x9d)), "(CMF<<8|FLG) % 31 != 0 (:1214-1220)")
    add("zlib_fdict", png_file(w, h, 8, 6, 0, zstream(rows4, fdict=True)), "FDICT set, check bits valid (:1262-1265)")
    add("zlib_cinfo_small_window", png_file(w, h, 8, 6, 0, zstream(rows4, cmf=0x28)), "CINFO = 2 (1 KiB window): only CM is looked at")
    grey = img[..., :1]
    add("colour_type_0_grey8", png_file(w, h, 8, 0, 0, zstream(corpus.png_filter_rows(grey, 0))), "colour type 0 (:987-1038)")
    ga = img[..., :2]
    add("colour_type_4_greyalpha", png_file(w, h, 8, 4, 0, zstream(corpus.png_filter_rows(ga, 1))), "colour type 4")
    wide = np.repeat(img, 2, axis=2)
    add("bit_depth_16_rgba", png_file(w, h, 16, 6, 0, zstream(corpus.png_filter_rows(wide, 0))), "bit depth 16 (:1088)")
    add("bit_depth_4_palette", png_file(w, h, 4, 3, 0, zstream(bytes((w // 2 + 1) * h)), extra=[(b"PLTE", bytes(48))]), "bit depth 4, colour type 3")
    add("filter_method_1", b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 1, 0)) + chunk(b"IDAT", zstream(rows4)) + chunk(b"IEND", b""),
        "filter method 1 (:1118)")
    add("compression_method_1", b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 1, 0, 0)) + chunk(b"IDAT", zstream(rows4)) + chunk(b"IEND", b""),
        "IHDR compression method 1: not looked at by the reference")
    add("interlace_flag_plain_data", png_file(w, h, 8, 6, 1, zstream(rows4)),
        "interlace = 1 over NON-interlaced data: the silent build never looks at the flag (:1115), so this decodes like the control")
    add("unknown_critical_chunk", png_file(w, h, 8, 6, 0, zstream(rows4), extra=[(b"ABCD", b"xyz")]), "unknown critical chunk (:1309-1319)")
    add("ancillary_chunks", png_file(w, h, 8, 6, 0, zstream(rows4), extra=[(b"tEXt", b"k\0v"), (b"gAMA", struct.pack(">I", 45455))]), "ancillary chunks are skipped (:1303)")
    add("zero_width", png_file(0, h, 8, 6, 0, zstream(bytes(h))), "w = 0 (:1044)")
    return out


def gz_vectors():
    out = []
    text = corpus.word_salad(4000, 21)
    d = corpus.raw_deflate(text, 6)
    tail = struct.pack("<II", zlib.crc32(text) & 0xFFFFFFFF, len(text))

    def member(flg, extra=b""):
        return bytes([31, 139, 8, flg, 0, 0, 0, 0, 0, 255]) + extra + d + tail

    cases = {
        "gz_fcomment": (member(0x10, b"a comment\0"), "FLG = 0x10: the silent build does not skip the comment (decode_gz.c:223-233)"),
        "gz_fname_fcomment": (member(0x18, b"name.txt\0a comment\0"), "FLG = 0x18: FNAME skipped, FCOMMENT not"),
        "gz_fextra": (member(0x04, struct.pack("<H", 4) + b"ABCD"), "FLG = 0x04 FEXTRA: ignored by the reference (Q9)"),
        "gz_fhcrc": (member(0x02, b"\x12\x34"), "FLG = 0x02 FHCRC: ignored by the reference (Q9)"),
        "gz_ftext_only": (member(0x01), "FLG = 0x01 FTEXT: no extra field, decodes"),
        "gz_fname_unterminated": (bytes([31, 139, 8, 8, 0, 0, 0, 0, 0, 255]) + b"abcdefgh" * 4, "FNAME without terminator"),
        "gz_too_short": (bytes([31, 139, 8, 0, 0, 0, 0, 0, 0, 255]) + b"\x03\x00" + bytes(4), "10-byte header + 6 bytes"),
    }
    for name, (g, note) in cases.items():
        good, o = reflib.decode_gz(g, 20000)
        out.append(dict(name=name, in_b64=b64(g), cap=20000, good=good, out_len=len(o), out_sha256=sha(o), note=note,
                        equals_source=bool(good and o == text)))
    return out


def fixtures():
    from PIL import Image
    out = {}
    for name in ("fs_angrymob.png", "fs_birdmystic.png", "fs_bridge.png", "fs_cannon.png"):
        d = open(os.path.join(HERE, name), "rb").read()
        good, w, h, rgba = reflib.decode_png(d)
        pil = Image.open(io.BytesIO(d)).convert("RGBA").tobytes()
        out[name] = dict(kind="png", good=good, w=w, h=h, ref_sha256=sha(rgba), spec_sha256=sha(pil), equals_spec=bool(rgba == pil), bytes=len(d))
    return out


if __name__ == "__main__":
    m = dict(generator="tests/golden/make_golden_r2.py (reference = oracle/_ref/libref.so, silent no-assert build)",
             fixtures=fixtures(), png=png_vectors(), gz=gz_vectors())
    with open(os.path.join(HERE, "manifest_r2.json"), "w") as f:
        json.dump(m, f, indent=1)
    for k, v in m["fixtures"].items():
        print(k, v["good"], v["equals_spec"], v["w"], v["h"])
    for v in m["png"]:
        print("png", v["name"], v["good"])
    for v in m["gz"]:
        print("gz ", v["name"], v["good"], v["out_len"], v["equals_source"])
