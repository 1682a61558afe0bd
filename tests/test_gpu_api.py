"""GPU tests of every entry point of the C-ABI: the reference-compatible scalar
API (include/inflate.h, decode_png.h, decode_gz.h + legacy aliases), the packed
host API (single- and multi-wave), and the device-resident API."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import debigulator_b200 as dbg
from debigulator_b200 import corpus

pytestmark = pytest.mark.gpu


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def L():
    return dbg.load_library()


def test_scalar_inflate(L, ctx, ref):
    L.inflate_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.inflate.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                          C.POINTER(C.c_uint32), C.c_uint32]
    L.inflate.restype = None
    L.inflate_init(None, None, None, 1)
    data = corpus.word_salad(50000, 3)
    for level, strat in ((6, 0), (6, 4), (0, 0)):
        s = corpus.raw_deflate(data, level, strat)
        cap = max(len(data), len(s)) + 64
        ib = C.create_string_buffer(s + bytes(16), len(s) + 16)
        ob = C.create_string_buffer(cap)
        n, g = C.c_uint64(123), C.c_uint32(7)
        L.inflate(ob, cap, C.byref(n), None, 0, ib, len(s), C.byref(g), 1)
        rg, rout = ref.inflate(s, cap)
        assert g.value == rg == 1 and ob.raw[: n.value] == rout == data
    # argument checks (inflate.c:797-844)
    g = C.c_uint32(7)
    n = C.c_uint64(0)
    L.inflate(None, 10, C.byref(n), None, 0, ib, 10, C.byref(g), 1)
    assert g.value == 0
    L.inflate(ob, 3, C.byref(n), None, 0, ib, 10, C.byref(g), 1)      # recipient_size < compressed_input_size
    assert g.value == 0
    L.inflate(ob, 100, C.byref(n), None, 0, ib, 4, C.byref(g), 1)     # compressed_input_size < 5
    assert g.value == 0


def test_scalar_png_and_legacy(L, golden_dir, manifest):
    L.decode_png_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.decode_png.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint8)]
    L.decode_png.restype = None
    L.decode_png_deinit.argtypes = [C.c_uint32]
    data = open(os.path.join(golden_dir, "gimp_test.png"), "rb").read()
    want = manifest["fixtures"]["gimp_test.png"]["ref_sha256"]
    ib = C.create_string_buffer(data, len(data))
    ob = C.create_string_buffer(1024 * 1024 * 4)
    g = C.c_uint8(9)
    L.decode_png(ib, len(data), ob, len(ob), 2, C.byref(g))            # slot 2 not initialised (decode_png.c:691)
    assert g.value == 0
    L.decode_png_init(None, None, None, None, 1024 * 1024 * 4 + 1024 + 1 + 3000000, 2)
    L.decode_png(ib, len(data), ob, len(ob), 2, C.byref(g))
    assert g.value == 1 and sha(ob.raw) == want
    assert ib.raw == data                                              # unlike the reference, the input is untouched
    L.decode_png(ib, len(data), ob, len(ob) - 4, 2, C.byref(g))        # rgba size mismatch (decode_png.c:970)
    assert g.value == 0
    L.decode_png_deinit(2)
    L.decode_png_init(None, None, None, None, 1000000, 2)              # working memory too small (:1060-1081)
    L.decode_png(ib, len(data), ob, len(ob), 2, C.byref(g))
    assert g.value == 0
    L.decode_png_deinit(2)
    # legacy names (hellopng.c:154-200)
    L.init_PNG_decoder.argtypes = [C.c_void_p]
    L.get_PNG_width_height.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.decode_PNG.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    L.init_PNG_decoder(None)
    w, h, g32 = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    L.get_PNG_width_height(ib, len(data), C.byref(w), C.byref(h), C.byref(g32))
    assert (g32.value, w.value, h.value) == (1, 1024, 1024)
    L.decode_PNG(ib, len(data), ob, len(ob), C.byref(g32))
    assert g32.value == 1 and sha(ob.raw) == want


class DecodedData(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_size", C.c_uint32), ("good", C.c_uint32)]


def test_scalar_gz(L, golden_dir, manifest):
    libc = C.CDLL(None)
    L.init_decode_gz.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.decode_gz.argtypes = [C.c_void_p, C.c_uint32]
    L.decode_gz.restype = C.POINTER(DecodedData)
    data = open(os.path.join(golden_dir, "gzipsample.gz"), "rb").read()
    ib = C.create_string_buffer(data, len(data))
    L.init_decode_gz(C.cast(libc.malloc, C.c_void_p), C.cast(libc.memset, C.c_void_p), C.cast(libc.memcpy, C.c_void_p))
    r = L.decode_gz(ib, len(data))
    f = manifest["fixtures"]["gzipsample.gz"]
    assert r.contents.good == 1 and r.contents.data_size == f["out_len"]
    assert sha(C.string_at(r.contents.data, r.contents.data_size)) == f["ref_sha256"]
    bad = C.create_string_buffer(b"\x1f\x8c" + data[2:], len(data))
    r2 = L.decode_gz(bad, len(data))
    assert r2.contents.good == 0 and not r2.contents.data and r2.contents.data_size == 0     # Q15 fixed


def test_packed_api_waves_and_ragged(ctx, ref):
    """Monotonic arenas take the multi-wave pipelined path; shuffled offsets the single-wave path.
    Expected payloads come from the reference (rule Q2 shortens some low-entropy members)."""
    n = 600
    base = [corpus.gz_member_cfg5(i, 20000 + 977 * (i % 13)) for i in range(40)]
    base = [(g, ref.decode_gz(g, len(d) + len(g))[1]) for g, d in base]
    items = [base[i % 40] for i in range(n)]
    for shuffled in (False, True):
        order = list(range(n))
        if shuffled:
            np.random.default_rng(1).shuffle(order)
        in_off, out_off = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        ti = to = 0
        for i in order:
            g, d = items[i]
            in_off[i], out_off[i] = ti, to
            ti += (len(g) + 16 + 15) // 16 * 16
            to += (len(d) + len(g) + 16 + 15) // 16 * 16
        h_in = np.zeros(ti + 64, np.uint8)
        h_out = np.zeros(to + 64, np.uint8)
        for i, (g, d) in enumerate(items):
            h_in[int(in_off[i]):int(in_off[i]) + len(g)] = np.frombuffer(g, np.uint8)
        in_size = np.array([len(g) for g, _ in items], np.uint64)
        out_cap = np.array([len(d) + len(g) + 16 for g, d in items], np.uint64)
        osz, st = ctx.decode_packed(dbg.api.KIND_GZ, h_in, in_off, in_size, h_out, out_off, out_cap)
        assert int(st.sum()) == 0
        for i, (g, d) in enumerate(items):
            assert int(osz[i]) == len(d)
            assert h_out[int(out_off[i]):int(out_off[i]) + len(d)].tobytes() == d, (shuffled, i)


def test_device_api_misaligned_inputs(ctx):
    import torch
    dev = torch.device("cuda", 0)
    data = [corpus.word_salad(30000 + 1000 * i, i) for i in range(33)]
    streams = [corpus.raw_deflate(d, 6 if i % 2 else 1) for i, d in enumerate(data)]
    offs, total = [], 16
    for i, s in enumerate(streams):
        total += (i * 7) % 16 + 1           # every alignment 0..15 occurs
        offs.append(total)
        total += len(s)
    arena = np.zeros((total + 64 + 15) // 16 * 16, np.uint8)
    for o, s in zip(offs, streams):
        arena[o:o + len(s)] = np.frombuffer(s, np.uint8)
    caps = [len(d) + 64 for d in data]
    out_off = np.cumsum([0] + caps[:-1]).astype(np.uint64)
    i64 = lambda a: torch.from_numpy(np.asarray(a, np.uint64).view(np.int64)).to(dev)
    d_in = torch.from_numpy(arena).to(dev)
    d_out = torch.zeros(int(sum(caps)), dtype=torch.uint8, device=dev)
    d_size = torch.zeros(len(data), dtype=torch.int64, device=dev)
    d_st = torch.full((len(data),), -1, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        ctx.inflate_device(d_in, i64(offs), i64([len(x) for x in streams]), d_out, i64(out_off), i64(caps), d_size, d_st,
                           order=None, stream=s.cuda_stream)
    s.synchronize()
    assert d_st.abs().sum().item() == 0
    out = d_out.cpu().numpy()
    for i, d in enumerate(data):
        assert int(d_size[i]) == len(d)
        assert out[int(out_off[i]):int(out_off[i]) + len(d)].tobytes() == d, i


def test_png_device_api_large_images(ctx, ref):
    """One 2048x1536 image per filter mode through the device-resident PNG entry point."""
    import torch
    dev = torch.device("cuda", 0)
    imgs = [corpus.gradient_noise_rgba(2048, 1536, 40 + i) for i in range(3)]
    files = [corpus.write_png(im, f) for im, f in zip(imgs, (4, 3, -1))]
    offs, total = [], 0
    for f in files:
        offs.append(total)
        total += (len(f) + 16 + 15) // 16 * 16
    arena = np.zeros(total + 64, np.uint8)
    for o, f in zip(offs, files):
        arena[o:o + len(f)] = np.frombuffer(f, np.uint8)
    rgba = 2048 * 1536 * 4
    i64 = lambda a: torch.from_numpy(np.asarray(a, np.uint64).view(np.int64)).to(dev)
    d_in = torch.from_numpy(arena).to(dev)
    d_out = torch.zeros(3 * rgba, dtype=torch.uint8, device=dev)
    d_st = torch.full((3,), -1, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        ctx.png_device(d_in, i64(offs), i64([len(f) for f in files]), d_out, i64([0, rgba, 2 * rgba]), i64([rgba] * 3), d_st,
                       sum(len(f) for f in files), 3 * rgba, stream=s.cuda_stream)
    s.synchronize()
    assert d_st.tolist() == [0, 0, 0]
    out = d_out.cpu().numpy()
    for i, im in enumerate(imgs):
        assert out[i * rgba:(i + 1) * rgba].tobytes() == im.tobytes(), i


def test_failed_items_do_not_poison_batch(ctx):
    good = corpus.gz_member_cfg2(2, 1 << 15)
    bad_magic = b"\x00" + good[0][1:]
    truncated = good[0][: len(good[0]) // 2]
    items = [good[0], bad_magic, good[0], truncated, b"", good[0]]
    res = ctx.decode_gz_batch(items, [len(good[1]) + len(good[0])] * len(items))
    assert [r[0] for r in res] == [1, 0, 1, res[3][0], 0, 1]
    for k in (0, 2, 5):
        assert res[k][1] == good[1]


def test_example_mains_print_the_readme_golden(golden_dir, tmp_path):
    """examples/hellopng.c, hellogz.c and hellobmp.c (the roles of the reference's stale examples) build with gcc against
    include/*.h and print the reference README's golden for gimp_test.png (README.md:41-47)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "debigulator_b200")
    outs = {}
    for name in ("hellopng", "hellogz", "hellobmp"):
        exe = str(tmp_path / name)
        subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", name + ".c"),
                               "-L" + lib, "-ldebigulator_b200", "-Wl,-rpath," + lib, "-o", exe])
        arg = os.path.join(golden_dir, {"hellopng": "gimp_test.png", "hellogz": "gzipsample.gz", "hellobmp": "fs_psychologist.bmp"}[name])
        outs[name] = subprocess.run([exe, arg], capture_output=True, text=True, timeout=300)
        assert outs[name].returncode == 0, outs[name].stdout + outs[name].stderr
    p = outs["hellopng"].stdout
    for line in ("bytes read from raw file: 30522", "result was: SUCCESS", "rgba values in image: 4194304",
                 "pixels in image (info from image header): 1048576", "image width: 1024", "image height: 1024",
                 "average pixel: [248,249,251,158]"):
        assert line in p, p
    assert "decompressed bytes: 561872" in outs["hellogz"].stdout
    b = outs["hellobmp"].stdout
    for line in ("image width: 406", "image height: 610", "decode_BMP result was: SUCCESS", "encode_BMP wrote: 990695 bytes",
                 "round trip: EXACT"):
        assert line in b, b


def test_optin_gzip_trailer_verification(ctx):
    """dbg_set_verify: CRC32 / ISIZE are checked on the device only when asked; by default `good` follows the
    reference, which ignores the trailer (decode_gz.c:281-297)."""
    import struct
    import zlib
    data = corpus.word_salad(200000, 21)
    ok = corpus.gzip_frame(corpus.raw_deflate(data), data)
    bad_crc = ok[:-8] + struct.pack("<II", (zlib.crc32(data) ^ 1) & 0xFFFFFFFF, len(data))
    bad_size = ok[:-8] + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) + 1)
    items = [ok, bad_crc, bad_size]
    caps = [len(data) + len(ok)] * 3
    assert [g for g, _ in ctx.decode_gz_batch(items, caps)] == [1, 1, 1]
    ctx.set_verify(True)
    try:
        res = ctx.decode_gz_batch(items, caps)
        assert [g for g, _ in res] == [1, 0, 0]
        assert res[0][1] == data
    finally:
        ctx.set_verify(False)


def test_optin_png_adler_verification(ctx):
    """With dbg_set_verify the zlib Adler-32 of the scanline stream is checked (single- and multi-IDAT files)."""
    import struct
    import zlib
    img = corpus.gradient_noise_rgba(200, 150, 31)
    ok1 = corpus.write_png(img, 4, single_block=True)
    ok2 = corpus.write_png(img, -1, idat_split=4000, strategy=zlib.Z_DEFAULT_STRATEGY)

    def break_adler(png):
        # flip a bit in the last 4 bytes of the (last) IDAT payload and fix that chunk's CRC
        pos, last = 8, None
        while pos < len(png):
            (ln,) = struct.unpack(">I", png[pos:pos + 4])
            if png[pos + 4:pos + 8] == b"IDAT":
                last = (pos, ln)
            pos += 12 + ln
        p, ln = last
        body = bytearray(png[p + 4:p + 8 + ln])
        body[-1] ^= 1
        return png[:p + 4] + bytes(body) + struct.pack(">I", zlib.crc32(bytes(body)) & 0xFFFFFFFF) + png[p + 12 + ln:]

    files = [ok1, ok2, break_adler(ok1), break_adler(ok2)]
    assert [r[0] for r in ctx.decode_png_batch(files)] == [1, 1, 1, 1]      # the reference ignores Adler-32
    ctx.set_verify(True)
    try:
        res = ctx.decode_png_batch(files)
        assert [r[0] for r in res] == [1, 1, 0, 0]
        assert res[0][3] == img.tobytes() and res[1][3] == img.tobytes()
    finally:
        ctx.set_verify(False)


# ------------------------------------------------------------------ round 2: waves, multi-device, stream order --
def _png_batch(n, w=192, h=160, unique=12):
    uniq = [corpus.png_cfg3(i, w, h) for i in range(unique)]
    files = [uniq[i % unique][0] for i in range(n)]
    want = [uniq[i % unique][1] for i in range(n)]
    return files, want


def _pack(files, caps, pad=16):
    in_off, out_off, ti, to = [], [], 0, 0
    for f, c in zip(files, caps):
        in_off.append(ti)
        ti += (len(f) + pad + 15) // 16 * 16
        out_off.append(to)
        to += (c + 15) // 16 * 16
    h_in = np.zeros(ti + 64, np.uint8)
    for f, o in zip(files, in_off):
        h_in[o:o + len(f)] = np.frombuffer(f, np.uint8)
    return h_in, np.array(in_off, np.uint64), np.array([len(f) for f in files], np.uint64), np.zeros(to + 64, np.uint8), \
        np.array(out_off, np.uint64), np.array(caps, np.uint64)


def test_packed_png_waves(ctx, monkeypatch):
    """PNG through the packed host API in several overlapped waves (per-wave scratch, per-wave lane-serial path)."""
    monkeypatch.setenv("DBG_PNG_WAVES", "8")
    c = dbg.Context(0)
    # > 64 MB of files so that the batch is really cut into waves
    files, want = _png_batch(96, 1024, 768, 6)
    files[5] = files[5][:-20]                      # truncated: IEND missing
    bad = bytearray(files[7])
    bad[len(bad) // 2] ^= 0x10
    files[7] = bytes(bad)                          # IDAT bit flip -> CRC mismatch
    caps = [1024 * 768 * 4] * len(files)
    h_in, in_off, in_size, h_out, out_off, out_cap = _pack(files, caps)
    assert int(in_size.sum()) > 2 * (64 << 20)
    fx0 = c.fx_stats()
    osz, st = c.decode_packed(dbg.api.KIND_PNG, h_in, in_off, in_size, h_out, out_off, out_cap)
    assert st[5] != 0 and st[7] != 0 and osz[5] == 0
    for i in range(len(files)):
        if i in (5, 7):
            continue
        assert st[i] == 0 and osz[i] == caps[i], (i, st[i])
        o = int(out_off[i])
        assert h_out[o:o + caps[i]].tobytes() == want[i], i
    assert c.fx_stats()[0] - fx0[0] >= 90          # the waves used the lane-serial path
    c.close()


def test_packed_multi_two_contexts_one_gpu(ref):
    """dbg_decode_batch_packed_multi with two contexts (both on cuda:0, so the partition, the per-device threads and the
    sub-set waves all run on a one-GPU box): gzip members and PNGs, every item against the reference."""
    m = dbg.MultiContext(2, [0, 0])
    assert m.n_devices == 2
    members = [corpus.gz_member_cfg2(i, 1 << 17) for i in range(600)]
    caps = [len(d) + len(g) + 64 for g, d in members]
    h_in, in_off, in_size, h_out, out_off, out_cap = _pack([g for g, _ in members], caps)
    osz, st, dev = m.decode_packed(dbg.api.KIND_GZ, h_in, in_off, in_size, h_out, out_off, out_cap)
    assert set(dev.tolist()) == {0, 1}
    assert abs(int((dev == 0).sum()) - 300) < 120
    for i, (g, d) in enumerate(members):
        assert st[i] == 0 and osz[i] == len(d), i
        o = int(out_off[i])
        assert h_out[o:o + len(d)].tobytes() == d, i
    rg, rout = ref.decode_gz(members[3][0], caps[3])
    assert rg == 1 and rout == members[3][1]
    files, want = _png_batch(40, 640, 480, 8)
    caps = [640 * 480 * 4] * len(files)
    h_in, in_off, in_size, h_out, out_off, out_cap = _pack(files, caps)
    osz, st, dev = m.decode_packed(dbg.api.KIND_PNG, h_in, in_off, in_size, h_out, out_off, out_cap)
    assert set(dev.tolist()) == {0, 1}
    for i in range(len(files)):
        assert st[i] == 0, i
        o = int(out_off[i])
        assert h_out[o:o + caps[i]].tobytes() == want[i], i
    assert m.fx_stats(0)[0] + m.fx_stats(1)[0] == 40
    m.close()


def test_pipe_keeps_batches_in_flight(ref):
    """dbg_pipe_*: five packed batches (gzip, PNG, gzip, ...) through a pipe of depth 2, tickets waited for out of order,
    every item against what the members / images were made from; one member against the reference."""
    p = dbg.Pipe(0, 2)
    assert p.depth == 2
    jobs = []
    for b in range(5):
        if b % 2 == 0:
            members = [corpus.gz_member_cfg2(7 * b + i, 1 << 16) for i in range(300 + 50 * b)]
            caps = [len(d) + len(g) + 64 for g, d in members]
            packed = _pack([g for g, _ in members], caps)
            want = [d for _, d in members]
            kind = dbg.api.KIND_GZ
            if b == 0:
                rg, rout = ref.decode_gz(members[5][0], caps[5])
                assert rg == 1 and rout == members[5][1]
        else:
            files, want = _png_batch(24, 640, 480, 8)
            caps = [640 * 480 * 4] * len(files)
            packed = _pack(files, caps)
            kind = dbg.api.KIND_PNG
        h_in, in_off, in_size, h_out, out_off, out_cap = packed
        jobs.append((p.submit(kind, h_in, in_off, in_size, h_out, out_off, out_cap), packed, want, kind))
    for k in (1, 0, 2, 4, 3):
        ticket, (h_in, in_off, in_size, h_out, out_off, out_cap), want, kind = jobs[k]
        osz, st = p.wait(ticket)
        for i, d in enumerate(want):
            assert st[i] == 0 and osz[i] == len(d), (k, i)
            o = int(out_off[i])
            assert h_out[o:o + len(d)].tobytes() == d, (k, i)
    assert p.kernel_launches() > 0
    assert p.L.dbg_pipe_wait(p.h, jobs[0][0]) == -3 and p.L.dbg_pipe_wait(p.h, 99) == -3  # waited for already / never handed out: DBG_ERR_ARG
    p.close()


def test_device_calls_on_two_streams_share_scratch_safely(ctx):
    """Two device-resident calls issued back to back on DIFFERENT streams, neither synchronised in between: the
    second must not disturb the first one's work queue and descriptors (per-context scratch)."""
    import torch
    dev = torch.device("cuda", 0)
    members = [corpus.gz_member_cfg2(i, 1 << 18) for i in range(64)]
    n = 512
    caps = [(1 << 18) + 4096] * n
    h_in, in_off, in_size, _, out_off, out_cap = _pack([members[i % 64][0] for i in range(n)], caps)
    i64 = lambda a: torch.from_numpy(np.asarray(a, np.uint64).view(np.int64)).to(dev)
    d_in = torch.from_numpy(h_in).to(dev)
    a_off, a_sz, o_off, o_cap = i64(in_off), i64(in_size), i64(out_off), i64(out_cap)
    outs, sizes, sts = [], [], []
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    for s in streams:
        outs.append(torch.zeros(int(out_off[-1] + out_cap[-1]) + 64, dtype=torch.uint8, device=dev))
        sizes.append(torch.zeros(n, dtype=torch.int64, device=dev))
        sts.append(torch.full((n,), 77, dtype=torch.int32, device=dev))
    for k, s in enumerate(streams):
        ctx.inflate_device(d_in, a_off, a_sz, outs[k], o_off, o_cap, sizes[k], sts[k], None, stream=s.cuda_stream, gz=True)
    torch.cuda.synchronize()
    for k in range(2):
        assert int(sts[k].abs().sum().item()) == 0, k
        assert bool((sizes[k] == (1 << 18)).all().item()), k
    assert torch.equal(outs[0], outs[1])
    got = outs[0].cpu().numpy()
    for i in (0, 1, 2, 3, 63, 511):
        o = int(out_off[i])
        assert got[o:o + (1 << 18)].tobytes() == members[i % 64][1], i


def test_scalar_gz_lying_isize_and_trim(L, ctx):
    """An 18-byte-header member whose ISIZE claims 4 GiB must not make the drop-in decode_gz() reserve that (the bound
    is what DEFLATE can expand to); dbg_trim releases the scratch and the context keeps working."""
    import struct
    import zlib
    L.init_decode_gz.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.decode_gz.argtypes = [C.c_void_p, C.c_uint32]

    class DD(C.Structure):
        _fields_ = [("data", C.c_void_p), ("data_size", C.c_uint32), ("good", C.c_uint32)]
    L.decode_gz.restype = C.POINTER(DD)
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    sizes = []
    MALLOC = C.CFUNCTYPE(C.c_void_p, C.c_size_t)

    def my_malloc(nbytes):
        sizes.append(nbytes)
        return libc.malloc(C.c_size_t(nbytes))
    cb = MALLOC(my_malloc)
    L.init_decode_gz(cb, None, None)
    data = corpus.word_salad(30000, 9)
    g = bytearray(corpus.gzip_frame(corpus.raw_deflate(data, 6), data))
    g[-4:] = struct.pack("<I", 0xfffffff0)        # ISIZE lies
    ib = C.create_string_buffer(bytes(g), len(g))
    r = L.decode_gz(ib, len(g))
    assert r.contents.good == 1 and r.contents.data_size == len(data)
    assert C.string_at(r.contents.data, len(data)) == data
    assert max(sizes) <= len(data) + 64           # the struct and an exact-size buffer, nothing speculative
    ctx.trim()
    (good, out), = ctx.decode_gz_batch([bytes(corpus.gzip_frame(corpus.raw_deflate(data, 6), data))], [len(data) + 64])
    assert good == 1 and out == data
