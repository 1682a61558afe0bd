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

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "simt_emu")
CSRC = os.path.join(os.path.dirname(HERE), "debigulator_b200", "csrc")


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libsimt_emu.so")
    srcs = [os.path.join(EMU_DIR, "emu.cpp")] + [os.path.join(CSRC, f) for f in ("simt.h", "inflate_core.h", "png_core.h", "bsplit_core.h", "fx_core.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", srcs[0], "-o", so])
    L = C.CDLL(so)
    L.emu_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int]
    L.emu_inflate.restype = C.c_uint32
    L.emu_png_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int]
    L.emu_png_decode.restype = C.c_uint32
    L.emu_fx_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_uint32,
                                 C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.emu_fx_inflate.restype = C.c_uint32
    L.emu_bsplit_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                     C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_int]
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


def test_block_split_token_pieces(emu, ref):
    """cut_token_pieces (bsplit_core.h): a chunk's token run expanded as several pieces, each a marker domain of its own --
    what a lone long stream gets so that its expansion is not one warp's latency chain. Same streams as the lane-parallel
    test, several piece counts, against the reference's inflate()."""
    import zlib
    from debigulator_b200 import corpus
    emu.emu_bsplit_inflate_pieces.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                              C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_int, C.c_uint32]
    emu.emu_bsplit_inflate_pieces.restype = C.c_uint32
    text = corpus.word_salad(300000, 33)
    img = corpus.gradient_noise_rgba(200, 150, 6)
    cases = [corpus.raw_deflate(text, 6), corpus.raw_deflate(text, 1), corpus.raw_deflate(corpus.png_filter_rows(img, 4), 6),
             corpus.raw_deflate(corpus.runs(600000, 7), 6), corpus.raw_deflate(corpus.periodic(500000, 3, 31000), 6)]
    pieced = 0
    for k, z in enumerate(cases):
        cap = 700000
        want_good, want = ref.inflate(z, cap)
        for region, pieces in ((32768, 16), (16384, 4), (65536, 7)):
            ib = C.create_string_buffer(z, len(z))
            ob = C.create_string_buffer(cap + 64)
            n, nch = C.c_uint64(0), C.c_uint32(0)
            st = emu.emu_bsplit_inflate_pieces(ib, len(z), ob, cap, C.byref(n), (5 * k) % 16, k & 1, region, C.byref(nch), 4, 1, pieces)
            if st == 0x4000:
                continue
            assert st == 0 and want_good == 1, (k, region, pieces, hex(st))
            assert ob.raw[: n.value] == want, (k, region, pieces)
            pieced += nch.value >> 16
    assert pieced >= 10


def test_paeth_swar_exhaustive(emu):
    """paeth4_swar (four channels per 32-bit word, png_core.h) equals the scalar Paeth predictor (decode_png.c:441-487) for
    all 2^24 byte triples. (The device build replaces two helpers by VABSDIFF4 / PRMT; the GPU parity suite covers those.)"""
    emu.emu_paeth4_mismatches.restype = C.c_uint64
    assert emu.emu_paeth4_mismatches() == 0


def test_png_unfilter_row_classes(emu, ref):
    """RGBA8 bands by row class (png_unfilter_band4): None/Up bands by columns, None/Sub bands by rows, the wavefront with and
    without Paeth rows, mixed bands, widths around the 32-pixel block and heights around the 32-row band."""
    from debigulator_b200 import corpus
    rng = np.random.default_rng(7)
    sets = [(0,), (1,), (2,), (3,), (4,), (0, 1), (0, 2), (1, 2), (0, 1, 2, 3), (0, 1, 2, 3, 4)]
    k = 0
    for w, h in ((1, 1), (1, 70), (5, 3), (31, 33), (32, 32), (33, 31), (64, 65), (97, 40), (130, 70)):
        for fs in sets:
            img = corpus.gradient_noise_rgba(w, h, 100 + k, amp=(3, 40)[k & 1])
            rows = rng.choice(fs, size=h)
            data = corpus.write_png(img, filt=rows, level=6, strategy=0)
            ob = C.create_string_buffer(w * h * 4 + 64)
            ib = C.create_string_buffer(data, len(data))
            good, _, _, want = ref.decode_png(data)  # tiny images fail in the reference (rule Q12): so must they here
            st = emu.emu_png_decode(ib, len(data), ob, w * h * 4, k & 1)
            assert (st == 0) == bool(good) and st < 0x1000, (w, h, fs, st)
            if good:  # the pixels are the spec's: the reference's own differ in the last rows of multi-block streams (defect D1)
                assert ob.raw[: w * h * 4] == img.tobytes(), (w, h, fs)
            k += 1


def _fx(emu, z, cap, mis, rev, chunk, group):
    ib = C.create_string_buffer(z, len(z))
    ob = C.create_string_buffer(cap + 64)
    n, ms, nc = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
    st = emu.emu_fx_inflate(ib, len(z), ob, cap, C.byref(n), mis, rev, chunk, group, C.byref(ms), C.byref(nc))
    return st, ob.raw[: n.value], ms.value, nc.value


def test_fx_lane_serial_kernel_source(emu, ref):
    """Lane-serial path for single fixed-Huffman-block streams (head -> sizes -> chain -> tokens -> expansion into
    16-bit cells -> resolve), the device source run by the emulator, against the reference's inflate()."""
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
    cases.append(bytes(200000))                                         # 258-byte matches at distance 1
    cases.append(bytes(rng.integers(0, 256, size=30000, dtype=np.uint8)))  # literals only (9-bit codes)
    ran = chunks = 0
    for k, data in enumerate(cases):
        for z in (corpus.fixed_block_deflate(data), None):
            if z is None:  # zlib's Z_FIXED, when it emits one block
                c = zlib.compressobj(9, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)
                z = c.compress(data) + c.flush()
            if (z[0] & 7) != 3 or len(z) < 4096:
                continue
            cap = max(len(data), len(z)) + 64
            want_good, want = ref.inflate(z, cap)
            assert want_good == 1 and want == data
            for chunk, group in ((2048, 1), (2048, 4), (4096, 16), (16384, 2)):
                st, out, ms, nc = _fx(emu, z, cap, (3 * k + chunk // 2048) % 16, (k + group) & 1, chunk, group)
                assert st == 0, (k, chunk, group, hex(st))
                assert out == want, (k, chunk, group)
                assert ms <= 3 or k in (2, 4, 5), (k, ms)  # strictly periodic symbol streams (runs, window-limit periods, literal-only) keep several chains alive
                ran += 1
                chunks += nc
    assert ran >= 24 and chunks > 300


def test_fx_lane_serial_damaged_streams(emu, ref):
    """Truncated streams (rule Q2 ends them, successfully), a cleared tail and flipped bits: the path reports what the
    reference's sequential decoder reports (status and bytes) whenever the reference has defined behaviour."""
    import numpy as np
    from debigulator_b200 import corpus
    img = corpus.gradient_noise_rgba(200, 150, 11)
    data = corpus.png_filter_rows(img, -1)
    z = corpus.fixed_block_deflate(data)
    cap = len(data) + 4096
    variants = [z[: len(z) // 2], z[: len(z) - 1], z[: 8192 + 5], z[:4096 + 1], z + bytes(5000)]
    rng = np.random.default_rng(3)
    for _ in range(12):
        b = bytearray(z)
        at = int(rng.integers(100, len(b) - 100))
        b[at] ^= 1 << int(rng.integers(0, 8))
        variants.append(bytes(b))
    checked = 0
    for k, v in enumerate(variants):
        if (v[0] & 7) != 3:
            continue
        want_good, want = ref.inflate(v, cap)
        st, out, _, _ = _fx(emu, v, cap, k % 16, k & 1, 2048, 3)
        assert st < 0x1000, (k, hex(st))
        if want_good:
            assert st == 0 and out == want, (k, st, len(out), len(want))
            checked += 1
        # where the reference fails (or runs into undefined behaviour) this path must simply fail or succeed cleanly
    assert checked >= 5


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
            st = emu.emu_bsplit_inflate(ib, len(z), ob, 400000, C.byref(n), (5 * k) % 16, k & 1, region, C.byref(nch), tpb, 0)
            assert st != 0x4000 and st < 0x1000, (k, tpb, hex(st))
            assert (st == 0) == bool(want_good), (k, tpb, st)
            if want_good:
                assert ob.raw[: n.value] == want, (k, tpb)
            used += nch.value & 0xffff
            expanded += nch.value >> 16
    assert used > 120 and expanded > 60  # the search really finds block boundaries, and tokens really get expanded


def test_lane_parallel_block_decode_kernel_source(emu, ref):
    """lane_decode_block (inflate_core.h): the count pass of the block-split path decodes Huffman blocks with one LANE per
    sub-chunk behind exact merge points. Streams with blocks of 10-40 KB (zlib's own block size), dynamic and fixed,
    text / PNG residuals / low entropy / runs; damaged copies must give the reference's verdict. The device source
    runs in the emulator, in both lane orders, against the reference's inflate()."""
    import zlib
    import numpy as np
    from debigulator_b200 import corpus
    emu.emu_lane_block_stats.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]
    tried, done = C.c_uint32(0), C.c_uint32(0)
    emu.emu_lane_block_stats(C.byref(tried), C.byref(done), 1)
    rng = np.random.default_rng(11)
    img = corpus.gradient_noise_rgba(256, 200, 5)
    text = corpus.word_salad(400000, 21)
    cases = [
        corpus.raw_deflate(text, 6),
        corpus.raw_deflate(text, 1),
        corpus.raw_deflate(text[:250000], 6, zlib.Z_FIXED),                       # multi-block fixed: no headers to find
        corpus.raw_deflate(corpus.png_filter_rows(img, 4), 6),
        corpus.raw_deflate(corpus.low_entropy(300000, 3, 5), 6, zlib.Z_HUFFMAN_ONLY),  # literals only, 2-3 bit codes
        corpus.raw_deflate(corpus.runs(2000000, 5), 6),                            # 258-byte matches: periodic symbol stream
        corpus.mixed_deflate(text, 4),
        corpus.raw_deflate(bytes(rng.integers(0, 256, 200000, dtype=np.uint8)) + text[:100000], 6),  # stored, then dynamic
    ]
    damaged = []
    for z in cases[:4]:
        b = bytearray(z)
        b[len(b) // 2 + 7] ^= 0x20
        damaged.append(bytes(b))
        damaged.append(z[: len(z) * 2 // 3])
    for k, z in enumerate(cases + damaged):
        cap = 2100000
        want_good, want = ref.inflate(z, cap)
        for region, tpb in ((32768, 4), (65536, 2)):
            ib = C.create_string_buffer(z, len(z))
            ob = C.create_string_buffer(cap + 64)
            n, nch = C.c_uint64(0), C.c_uint32(0)
            st = emu.emu_bsplit_inflate(ib, len(z), ob, cap, C.byref(n), (7 * k) % 16, k & 1, region, C.byref(nch), tpb, 1)
            if st == 0x4000:   # a hint that is no block boundary (damaged streams; stored text in the mixed one): the product
                assert k >= len(cases) or k == 6   # hands such a stream back to the warp-per-stream kernel
                continue
            assert st < 0x1000, (k, region, hex(st))
            if k < len(cases):
                assert st == 0 and want_good == 1
            if want_good and st == 0:
                assert ob.raw[: n.value] == want, (k, region)
            elif want_good:
                assert False, (k, region, st)
    emu.emu_lane_block_stats(C.byref(tried), C.byref(done), 0)
    assert done.value >= 60 and done.value >= tried.value // 2, (tried.value, done.value)


def test_lane_parallel_rounds_in_the_warp_per_stream_decoder(emu, ref):
    """inflate_warp (the warp-per-stream kernel's body) with its lane-parallel rounds: Huffman blocks are taken in rounds of
    8 KiB, tokens through the warp's scratch, expanded into the output at once; the ordinary symbol walk finishes what a
    round leaves. Whole streams, damaged streams, tight capacities -- every status and byte against the reference."""
    import zlib
    import numpy as np
    from debigulator_b200 import corpus
    emu.emu_lane_block_stats.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]
    tried, done = C.c_uint32(0), C.c_uint32(0)
    emu.emu_lane_block_stats(C.byref(tried), C.byref(done), 1)
    rng = np.random.default_rng(12)
    img = corpus.gradient_noise_rgba(200, 160, 6)
    text = corpus.word_salad(300000, 22)
    cases = [
        corpus.raw_deflate(text, 6),
        corpus.raw_deflate(text[:200000], 6, zlib.Z_FIXED),
        corpus.fixed_block_deflate(corpus.png_filter_rows(img, 4)),
        corpus.raw_deflate(corpus.low_entropy(200000, 3, 5), 6, zlib.Z_HUFFMAN_ONLY),
        corpus.raw_deflate(corpus.runs(1500000, 5), 6),
        corpus.raw_deflate(corpus.periodic(300000, 2, 31000), 9),
        corpus.mixed_deflate(text, 4),
        corpus.raw_deflate(bytes(rng.integers(0, 256, 100000, dtype=np.uint8)) + text[:100000], 6),
    ]
    variants = [(z, None) for z in cases]
    for z in cases[:4]:
        b = bytearray(z)
        b[len(b) // 2 + 3] ^= 0x04
        variants.append((bytes(b), None))
        variants.append((z[: len(z) * 3 // 5], None))
    want0 = ref.inflate(cases[0], 400000)[1]
    variants.append((cases[0], len(want0)))            # exact capacity
    variants.append((cases[0], len(want0) - 1))        # one byte short: overflow
    variants.append((cases[0], len(want0) // 2))
    for k, (z, cap) in enumerate(variants):
        tight = cap is not None
        cap = cap if tight else 1600000
        if cap < len(z):
            continue
        # (the reference writes past a capacity that is too small -- undefined behaviour, it is not called there)
        want_good, want = (1, want0) if tight else ref.inflate(z, cap)
        for rounds_off in (0, 16):
            st, out = emu_inflate(emu, z, cap, ((3 * k) % 16) | rounds_off, k & 1)
            assert st < 0x1000, (k, hex(st))
            if tight:
                assert (st == 0 and out == want0) if cap >= len(want0) else st == 9, (k, cap, st)
            elif k < len(cases):
                assert st == 0 and want_good == 1 and out == want, (k, rounds_off, st)
            elif want_good:
                # the reference succeeded on a damaged / short stream: same bytes, unless it ran over the capacity
                # (undefined behaviour there; this decoder reports the overflow)
                assert (st == 0 and out == want) or (st == 9 and len(want) > cap), (k, rounds_off, st)
            elif st == 0:
                assert False, (k, rounds_off, "reference failed, decoder did not")
    emu.emu_lane_block_stats(C.byref(tried), C.byref(done), 0)
    assert done.value >= 60, (tried.value, done.value)
