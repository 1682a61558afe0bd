"""CPU suite, part 2: the DEVICE kernel bodies (debigulator_b200/csrc/*_core.h)
compiled for the host by the 32-lane SIMT emulator in tests/simt_emu and checked
against the reference's golden vectors. This exercises the exact source the GPU
runs (bit reader, table builder, LZ77 copies, CRC combine, chunk walk,
wavefront un-filter) without a GPU. The emulator is test scaffolding only."""
import base64
import ctypes as C
import hashlib
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "simt_emu")
CSRC = os.path.join(os.path.dirname(HERE), "debigulator_b200", "csrc")


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libsimt_emu.so")
    srcs = [os.path.join(EMU_DIR, "emu.cpp")] + [os.path.join(CSRC, f) for f in ("simt.h", "inflate_core.h", "png_core.h", "bsplit_core.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", srcs[0], "-o", so])
    L = C.CDLL(so)
    L.emu_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int]
    L.emu_inflate.restype = C.c_uint32
    L.emu_png_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int]
    L.emu_png_decode.restype = C.c_uint32
    L.emu_split_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_uint32]
    L.emu_split_inflate.restype = C.c_uint32
    L.emu_bsplit_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                     C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]
    L.emu_bsplit_inflate.restype = C.c_uint32
    L.emu_crc32.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    L.emu_crc32.restype = C.c_uint32
    return L


def emu_inflate(L, data, cap, mis, rev):
    ib = C.create_string_buffer(data, max(len(data), 1))
    ob = C.create_string_buffer(cap + 64)
    n = C.c_uint64(0)
    st = L.emu_inflate(ib, len(data), ob, cap, C.byref(n), mis, rev)
    return st, ob.raw[: n.value]


def test_inflate_kernel_source_vs_golden(emu, manifest):
    bad = []
    for k, v in enumerate(manifest["inflate"]):
        if v["out_len"] > 120000:
            continue
        st, out = emu_inflate(emu, base64.b64decode(v["in_b64"]), v["cap"], (5 * k) % 16, k & 1)
        good = 1 if st == 0 else 0
        if st >= 0x1000 or good != v["good"] or (good and (len(out) != v["out_len"] or sha(out) != v["out_sha256"])):
            bad.append((v["name"], st))
    assert not bad


def test_crc_kernel_source(emu):
    import zlib
    for n in (1, 3, 4, 5, 17, 255, 256, 257, 4099, 8191, 8192, 8193, 20000, 65537):
        for off in (0, 1, 2, 3):
            d = os.urandom(n + off)
            b = C.create_string_buffer(d, len(d))
            assert emu.emu_crc32(C.addressof(b) + off, n, n & 1) == zlib.crc32(d[off:]), (n, off)


def test_png_kernel_source_vs_golden(emu, manifest):
    bad = []
    for k, v in enumerate(manifest["png"]):
        data = base64.b64decode(v["in_b64"])
        w = int.from_bytes(data[16:20], "big") if len(data) >= 24 else 0
        h = int.from_bytes(data[20:24], "big") if len(data) >= 24 else 0
        if w * h > 100 * 100:
            continue
        ob = C.create_string_buffer(w * h * 4 + 64)
        ib = C.create_string_buffer(data, len(data))
        st = emu.emu_png_decode(ib, len(data), ob, w * h * 4, k & 1)
        good = 1 if st == 0 else 0
        if st >= 0x1000 or good != v["good"] or (good and sha(ob.raw[: w * h * 4]) != v["out_sha256"]):
            bad.append((v["name"], st, v["good"]))
    assert not bad


def test_png_fixture_small(emu, manifest, golden_dir):
    for name in ("structuredart1.png", "structuredart2.png", "structuredart3.png", "font.png"):
        f = manifest["fixtures"][name]
        data = open(os.path.join(golden_dir, name), "rb").read()
        ob = C.create_string_buffer(f["w"] * f["h"] * 4 + 64)
        ib = C.create_string_buffer(data, len(data))
        assert emu.emu_png_decode(ib, len(data), ob, f["w"] * f["h"] * 4, 0) == 0
        assert sha(ob.raw[: f["w"] * f["h"] * 4]) == f["ref_sha256"], name


def test_split_stream_kernel_source(emu):
    """Split-stream path (transfer tables -> chain -> 16-bit cells -> resolve) on single fixed-Huffman
    block streams, against zlib. Streams come from zlib's Z_FIXED with one block (small inputs)."""
    import zlib
    import numpy as np
    from debigulator_b200 import corpus
    rng = np.random.default_rng(7)
    cases = []
    img = corpus.gradient_noise_rgba(160, 120, 3)
    cases.append(corpus.png_filter_rows(img, 4))                        # PNG-like residuals
    cases.append(corpus.word_salad(70000, 5))                           # text, long distances
    cases.append(corpus.periodic(120000, 2, 31000))                     # matches at the window limit
    cases.append(bytes(rng.integers(0, 4, size=90000, dtype=np.uint8))) # low entropy, short distances
    for k, data in enumerate(cases):
        # one fixed block: compress with Z_FIXED and keep only inputs zlib emits as a single block
        c = zlib.compressobj(9, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)
        z = c.compress(data) + c.flush()
        if (z[0] & 7) != 3:
            continue
        cap = len(data) + 64
        ib = C.create_string_buffer(z, len(z))
        ob = C.create_string_buffer(cap + 64)
        n = C.c_uint64(0)
        st = emu.emu_split_inflate(ib, len(z), ob, cap, C.byref(n), (3 * k) % 16, k & 1, (32768, 4096, 65536, 8192)[k % 4])
        assert st == 0, (k, st)
        assert ob.raw[: n.value] == zlib.decompress(z, -15), k


def test_block_split_kernel_source(emu, ref):
    """Block-split path (header search -> count -> chain -> 16-bit cells -> resolve) on multi-block streams,
    against the reference. memLevel 1 makes zlib close a block every 128 symbols, so small inputs already
    have hundreds of block boundaries; the mixed stream adds stored and fixed blocks between the dynamic ones."""
    import zlib
    import numpy as np
    from debigulator_b200 import corpus
    rng = np.random.default_rng(9)

    def many_blocks(data, level=6, mem=1):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, mem)
        return c.compress(data) + c.flush()

    cases = [
        (many_blocks(corpus.word_salad(60000, 1)), 2048),
        (many_blocks(corpus.periodic(150000, 2, 30000), 9, 2), 1024),
        (many_blocks(corpus.low_entropy(120000, 3)), 1500),
        (corpus.mixed_deflate(corpus.word_salad(200000, 4), 4), 4096),
        (many_blocks(corpus.runs(300000, 5)), 512),
        (corpus.raw_deflate(bytes(rng.integers(0, 256, 70000, dtype=np.uint8)), 6), 4096),  # stored blocks only: no hints
        (many_blocks(corpus.word_salad(60000, 6))[:-700], 2048),                             # truncated
    ]
    used = expanded = 0
    for k, (z, region) in enumerate(cases):
        want_good, want = ref.inflate(z, 400000)
        # tokens per compressed byte: 0 = both passes decode Huffman codes, 4 = the second pass expands the tokens
        # the first one recorded, 1 = too few slots for most chunks (mixes both)
        for tpb in (0, 4, 1):
            ib = C.create_string_buffer(z, len(z))
            ob = C.create_string_buffer(400000 + 64)
            n, nch = C.c_uint64(0), C.c_uint32(0)
            st = emu.emu_bsplit_inflate(ib, len(z), ob, 400000, C.byref(n), (5 * k) % 16, k & 1, region, C.byref(nch), tpb)
            assert st != 0x4000 and st < 0x1000, (k, tpb, hex(st))
            assert (st == 0) == bool(want_good), (k, tpb, st)
            if want_good:
                assert ob.raw[: n.value] == want, (k, tpb)
            used += nch.value & 0xffff
            expanded += nch.value >> 16
    assert used > 120 and expanded > 60  # the search really finds block boundaries, and tokens really get expanded
