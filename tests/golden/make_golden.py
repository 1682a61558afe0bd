"""Regenerates tests/golden/manifest.json by running the UNMODIFIED reference
(oracle/_ref/libref.so, built from /root/reference by oracle/Makefile) on
  * the bundled fixtures copied from /root/reference/resources (data, not code),
  * small synthetic deflate / gzip streams incl. the Appendix-A edge vectors,
  * small PNGs written by the reference's vendored stb_write.h (all filters).
Every entry records the reference's `good`, output size and sha256, and whether
that output equals the spec decoders (zlib / PIL), so a test can tell a
reference quirk from a bug. Run in the build container only:

    python tests/golden/make_golden.py
"""
import base64
import gzip
import hashlib
import io
import json
import os
import random
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reflib  # noqa: E402
from debigulator_b200 import corpus  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def b64(b):
    return base64.b64encode(b).decode()


def raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    return corpus.raw_deflate(data, level, strategy)


class BitWriter:
    def __init__(self):
        self.acc = 0
        self.n = 0
        self.out = bytearray()

    def bits(self, v, n):  # LSB-first
        self.acc |= (v & ((1 << n) - 1)) << self.n
        self.n += n
        while self.n >= 8:
            self.out.append(self.acc & 255)
            self.acc >>= 8
            self.n -= 8

    def code(self, c, n):  # Huffman code, MSB-first
        for k in range(n - 1, -1, -1):
            self.bits((c >> k) & 1, 1)

    def done(self):
        if self.n:
            self.out.append(self.acc & 255)
        return bytes(self.out)


def canon(lens):
    mx = max(lens)
    cnt = [0] * (mx + 2)
    for l in lens:
        if l:
            cnt[l] += 1
    code = 0
    nxt = [0] * (mx + 2)
    for b in range(1, mx + 1):
        code = (code + cnt[b - 1]) << 1
        nxt[b] = code
    out = {}
    for s, l in enumerate(lens):
        if l:
            out[s] = (nxt[l], l)
            nxt[l] += 1
    return out


def dynamic_block(litlen_lens, dist_lens, symbols, final=1):
    """Hand-rolled dynamic block: code lengths sent verbatim (no 16/17/18), all
    code-length codes 5 bits. symbols: list of ('lit', v) | ('eob',) |
    ('match', len_sym, len_extra_bits, len_extra, dist_sym, dist_xbits, dist_extra)."""
    w = BitWriter()
    w.bits(final, 1)
    w.bits(2, 2)
    hlit, hdist = len(litlen_lens), len(dist_lens)
    w.bits(hlit - 257, 5)
    w.bits(hdist - 1, 5)
    w.bits(19 - 4, 4)
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    cl = [5 if s < 16 else 0 for s in range(19)]  # 16 symbols x 5 bits = complete code
    for s in order:
        w.bits(cl[s], 3)
    clc = canon(cl)
    for l in list(litlen_lens) + list(dist_lens):
        w.code(*clc[l])
    lc, dc = canon(litlen_lens), canon(dist_lens) if any(dist_lens) else {}
    for s in symbols:
        if s[0] == 'lit':
            w.code(*lc[s[1]])
        elif s[0] == 'eob':
            w.code(*lc[256])
        else:
            _, ls, lxb, lx, ds, dxb, dx = s
            w.code(*lc[ls])
            w.bits(lx, lxb)
            w.code(*dc[ds])
            w.bits(dx, dxb)
    return w.done()


def inflate_vectors():
    vecs = []
    rnd = random.Random(1234)

    def add(name, stream, cap=None, data=None):
        if cap is None:
            cap = max(len(stream), len(data) if data is not None else 0) + 8
        good, out = reflib.inflate(stream, cap)
        spec = None
        try:
            spec = zlib.decompress(stream, -15)
        except Exception:
            spec = None
        vecs.append(dict(name=name, in_b64=b64(stream), cap=cap, good=good, out_len=len(out), out_sha256=sha(out),
                         equals_zlib=bool(good and spec is not None and out == spec)))

    text = corpus.word_salad(6000, 7)
    add("text_dynamic_l6", raw(text), data=text)
    add("text_dynamic_l1", raw(text, 1), data=text)
    add("text_dynamic_l9", raw(text, 9), data=text)
    add("text_fixed", raw(text, 6, zlib.Z_FIXED), data=text)
    add("text_huffman_only", raw(text, 6, zlib.Z_HUFFMAN_ONLY), data=text)
    add("stored_random", raw(os.urandom(3000), 0), data=bytes(3000))
    add("stored_two_blocks", raw(bytes(rnd.randrange(256) for _ in range(70000)), 0), data=bytes(70000))
    add("zeros_3098_q2", raw(bytes(3098)), data=bytes(3098))
    add("zeros_100000", raw(bytes(100000), 9), data=bytes(100000))
    add("rle_abc", raw(b"abc" * 3000, 6, zlib.Z_RLE), data=bytes(9000))
    add("mixed_full_flush", corpus.mixed_deflate(corpus.word_salad(20000, 9), 9), data=bytes(20000))
    for per in (1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 63, 64, 65, 257, 258, 259, 300, 1000, 32767, 32768):
        d = corpus.periodic(max(4 * per, 3000) if per < 20000 else 70000, per, per)
        add(f"period_{per}", raw(d, 9), data=d)
    for k in range(40):  # Q2 sweep: short low-entropy dynamic / huffman-only streams
        n = rnd.randrange(1, 400)
        kind = k % 4
        d = bytes(rnd.choice(b"ab") for _ in range(n)) if kind == 0 else bytes(n) if kind == 1 else \
            bytes(rnd.choice(b"abcd") for _ in range(n)) if kind == 2 else os.urandom(n)
        strat = rnd.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
        add(f"q2_sweep_{k}", raw(d, rnd.choice([1, 6, 9]), strat), data=d)
    # argument checks (inflate.c:826-844)
    s = raw(text)
    add("cap_lt_input", raw(os.urandom(2000), 0), cap=1500)
    add("input_lt_5", raw(b"a", 6, zlib.Z_FIXED)[:4], cap=100)
    # stored LEN/NLEN mismatch (inflate.c:949)
    st = bytearray(raw(os.urandom(300), 0))
    st[3] ^= 0x55
    add("stored_len_mismatch", bytes(st), cap=1000)
    # distance before start of output (inflate.c:1843)
    w = BitWriter(); w.bits(1, 1); w.bits(1, 2); w.code(0x30 + 65, 8); w.code(1, 7); w.bits(0b00100, 5); w.bits(1, 1); w.code(0, 7)
    add("distance_too_far", w.done() + bytes(4), cap=100)
    # hand-rolled dynamic blocks: HDIST=2 all-literal; 15-bit codes; lone dist code (Q3)
    ll = [0] * 257
    ll[97] = 1; ll[256] = 1
    add("dyn_all_literal_hdist2", dynamic_block(ll, [1, 1], [('lit', 97)] * 40 + [('eob',)]) + bytes(2), cap=200)
    add("dyn_lone_dist_q3", dynamic_block(ll, [1], [('lit', 97)] * 40 + [('eob',)]) + bytes(2), cap=200)
    # long codes: lengths 1,2,...,14,15,15 over 16 symbols (uses the 13-15 bit overflow path)
    ll = [0] * 257
    syms = list(range(65, 80)) + [256]
    for k, s_ in enumerate(syms):
        ll[s_] = min(k + 1, 15)
    body = [('lit', s_) for s_ in syms[:-1]] * 3 + [('eob',)]
    add("dyn_codes_to_15_bits", dynamic_block(ll, [1, 1], body) + bytes(2), cap=400)
    # matches with every length code / many distance codes through a dynamic table
    data = bytes(rnd.randrange(256) for _ in range(40000))
    add("random_40000_l9", raw(data + data[:20000] + data[100:3000], 9), data=bytes(70000))
    return vecs


def gz_vectors():
    out = []
    text = corpus.word_salad(5000, 11)
    cases = {
        "gz_flg0": corpus.gzip_frame(raw(text), text),
        "gz_fname": corpus.gzip_frame(raw(text), text, fname=b"hello.txt"),
        "gz_bad_magic": b"\x1f\x8c" + corpus.gzip_frame(raw(text), text)[2:],
        "gz_bad_cm": b"\x1f\x8b\x07" + corpus.gzip_frame(raw(text), text)[3:],
        "gz_python": gzip.compress(text, 6, mtime=0),
    }
    for name, g in cases.items():
        good, o = reflib.decode_gz(g, 20000)
        out.append(dict(name=name, in_b64=b64(g), cap=20000, good=good, out_len=len(o), out_sha256=sha(o)))
    return out


def png_vectors():
    out = []
    for (w, h) in [(1, 1), (1, 7), (7, 1), (5, 3), (33, 33), (64, 40), (100, 37)]:
        img = corpus.gradient_noise_rgba(w, h, w * 1000 + h, amp=9)
        for f in range(-1, 5):
            p = reflib.stb_png(img.tobytes(), w, h, 4, f)
            good, rw, rh, o = reflib.decode_png(p)
            out.append(dict(name=f"stb_{w}x{h}_f{f}", in_b64=b64(p), good=good, w=rw, h=rh, out_sha256=sha(o),
                            equals_source=bool(o == img.tobytes())))
    # incompressible image -> stb falls back to stored blocks -> reference rejects (Q12)
    noise = np.random.default_rng(5).integers(0, 256, size=(40, 40, 4), dtype=np.uint8)
    p = reflib.stb_png(noise.tobytes(), 40, 40, 4, 0)
    good, rw, rh, o = reflib.decode_png(p)
    out.append(dict(name="stb_noise_q12", in_b64=b64(p), good=good, w=rw, h=rh, out_sha256=sha(o), equals_source=False))
    # CRC flip, bad signature, truncated file
    img = corpus.gradient_noise_rgba(20, 20, 3)
    p = bytearray(reflib.stb_png(img.tobytes(), 20, 20, 4, 4))
    q = bytearray(p); q[-6] ^= 1   # IEND crc
    r = bytearray(p); r[50] ^= 0x10  # inside IDAT data
    s_ = bytearray(p); s_[1] = ord('Q')
    for name, v in (("png_iend_crc_flip", q), ("png_idat_bitflip", r), ("png_bad_signature", s_), ("png_truncated", p[:len(p) - 20])):
        good, rw, rh, o = reflib.decode_png(bytes(v))
        out.append(dict(name=name, in_b64=b64(bytes(v)), good=good, w=rw, h=rh, out_sha256=sha(o), equals_source=False))
    return out


def fixtures():
    from PIL import Image
    out = {}
    for name in sorted(os.listdir(HERE)):
        path = os.path.join(HERE, name)
        if name.endswith(".png"):
            d = open(path, "rb").read()
            good, w, h, rgba = reflib.decode_png(d)
            pil = Image.open(io.BytesIO(d)).convert("RGBA").tobytes()
            out[name] = dict(kind="png", good=good, w=w, h=h, ref_sha256=sha(rgba), spec_sha256=sha(pil),
                             equals_spec=bool(rgba == pil), bytes=len(d))
        elif name.endswith(".gz"):
            d = open(path, "rb").read()
            spec = gzip.decompress(d)
            good, o = reflib.decode_gz(d, len(spec) + len(d))
            out[name] = dict(kind="gz", good=good, out_len=len(o), ref_sha256=sha(o), spec_sha256=sha(spec),
                             equals_spec=bool(o == spec), bytes=len(d))
    return out


if __name__ == "__main__":
    m = dict(
        generator="tests/golden/make_golden.py (reference = oracle/_ref/libref.so, silent no-assert build)",
        fixtures=fixtures(), inflate=inflate_vectors(), gz=gz_vectors(), png=png_vectors())
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(m, f, indent=1)
    print("fixtures", len(m["fixtures"]), "inflate", len(m["inflate"]), "gz", len(m["gz"]), "png", len(m["png"]))
    for k, v in m["fixtures"].items():
        print(k, v["good"], v["equals_spec"])
    print("inflate vectors where ref != zlib:", [v["name"] for v in m["inflate"] if v["good"] and not v["equals_zlib"]])
    print("inflate vectors ref fails:", [v["name"] for v in m["inflate"] if not v["good"]])
    print("png ref fails:", [v["name"] for v in m["png"] if not v["good"]])
    print("gz:", [(v["name"], v["good"]) for v in m["gz"]])
