"""GPU parity tests: the CUDA path, called through the C-ABI, against
(a) the committed golden vectors produced by the unmodified reference and
(b) the reference itself (oracle/_ref) on seeded synthetic inputs.
Bit-exact on [0, out_size) and equal `good` flags."""
import base64
import hashlib
import os
import zlib

import numpy as np
import pytest

from debigulator_b200 import corpus

pytestmark = pytest.mark.gpu


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_inflate_golden_vectors(ctx, manifest):
    vecs = manifest["inflate"]
    res = ctx.inflate_batch([base64.b64decode(v["in_b64"]) for v in vecs], [v["cap"] for v in vecs])
    bad = []
    for v, (good, out) in zip(vecs, res):
        if good != v["good"] or (good and (len(out) != v["out_len"] or sha(out) != v["out_sha256"])):
            bad.append((v["name"], good, len(out), v["good"], v["out_len"]))
    assert not bad, bad


def test_gz_golden_vectors(ctx, manifest):
    vecs = manifest["gz"]
    res = ctx.decode_gz_batch([base64.b64decode(v["in_b64"]) for v in vecs], [v["cap"] for v in vecs])
    for v, (good, out) in zip(vecs, res):
        assert good == v["good"], v["name"]
        if good:
            assert len(out) == v["out_len"] and sha(out) == v["out_sha256"], v["name"]


def test_png_golden_vectors(ctx, manifest):
    vecs = manifest["png"]
    res = ctx.decode_png_batch([base64.b64decode(v["in_b64"]) for v in vecs])
    bad = []
    for v, (good, w, h, rgba) in zip(vecs, res):
        if good != v["good"] or (good and ((w, h) != (v["w"], v["h"]) or sha(rgba) != v["out_sha256"])):
            bad.append((v["name"], good, v["good"]))
    assert not bad, bad


def test_fixture_files(ctx, manifest, golden_dir):
    """BASELINE config 1 (gimp_test.png) and the other bundled fixtures.
    phoebus.png (D1) and backgrounddetailed1.png (D3) are the two documented
    divergences: there this decoder must match the spec decoders instead."""
    names = sorted(manifest["fixtures"])
    pngs = [n for n in names if manifest["fixtures"][n]["kind"] == "png"]
    res = ctx.decode_png_batch([open(os.path.join(golden_dir, n), "rb").read() for n in pngs])
    for n, (good, w, h, rgba) in zip(pngs, res):
        f = manifest["fixtures"][n]
        assert good == 1 and (w, h) == (f["w"], f["h"]), n
        want = f["ref_sha256"] if f["equals_spec"] else f["spec_sha256"]
        assert sha(rgba) == want, n
    gz = open(os.path.join(golden_dir, "gzipsample.gz"), "rb").read()
    f = manifest["fixtures"]["gzipsample.gz"]
    (good, out), = ctx.decode_gz_batch([gz], [f["out_len"] + len(gz)])
    assert good == 1 and len(out) == f["out_len"] and sha(out) == f["ref_sha256"]


def test_gimp_readme_golden(ctx, golden_dir):
    """README.md:41-47: 1024x1024, 4,194,304 RGBA values, average pixel [248,249,251,158]."""
    (good, w, h, rgba), = ctx.decode_png_batch([open(os.path.join(golden_dir, "gimp_test.png"), "rb").read()])
    assert good == 1 and (w, h) == (1024, 1024) and len(rgba) == 4194304
    avg = np.frombuffer(rgba, np.uint8).reshape(-1, 4).astype(np.uint64).sum(axis=0) // (w * h)
    assert list(avg) == [248, 249, 251, 158]


def test_gz_cfg2_members_vs_reference(ctx, ref):
    """BASELINE config 2 shape: 1 MiB members, classes stored/fixed/dynamic/mixed."""
    members = [corpus.gz_member_cfg2(i) for i in range(16)]
    caps = [len(d) + len(g) for g, d in members]
    res = ctx.decode_gz_batch([g for g, _ in members], caps)
    for i, ((g, d), (good, out)) in enumerate(zip(members, res)):
        rgood, rout = ref.decode_gz(g, caps[i])
        assert good == rgood == 1, i
        assert out == rout, i
        assert out == d, i


def test_cfg5_sweep_vs_reference(ctx, ref):
    sizes = [65536, 100000, 300000, 1 << 20, 70000, 2 << 20, 150000, 500000]
    members = [corpus.gz_member_cfg5(i, sizes[i % 8]) for i in range(16)]
    caps = [len(d) + len(g) + 64 for g, d in members]
    res = ctx.decode_gz_batch([g for g, _ in members], caps)
    for i, ((g, d), (good, out)) in enumerate(zip(members, res)):
        rgood, rout = ref.decode_gz(g, caps[i])
        assert good == rgood, i
        assert out == rout, i


def test_png_cfg3_vs_reference(ctx, ref):
    """BASELINE config 3 shape at reduced size plus two full-size images; all six filter modes,
    written by the reference's stb_write (one fixed-Huffman block) and by corpus.write_png."""
    files, want = [], []
    for i in range(12):
        img = corpus.gradient_noise_rgba(256, 192, 77 + i)
        files.append(ref.stb_png(img.tobytes(), 256, 192, 4, i % 6 - 1))
        want.append(img.tobytes())
    for i in range(2):
        img = corpus.gradient_noise_rgba(1024, 1024, 99 + i)
        files.append(ref.stb_png(img.tobytes(), 1024, 1024, 4, 4 if i else -1))
        want.append(img.tobytes())
    for i in range(6):
        p, rgba = corpus.png_cfg3(i, 300, 200)
        files.append(p)
        want.append(rgba)
    res = ctx.decode_png_batch(files)
    for i, (f, (good, w, h, rgba)) in enumerate(zip(files, res)):
        rgood, rw, rh, rrgba = ref.decode_png(f)
        assert good == rgood, i
        if good:
            assert rgba == rrgba, i
            assert rgba == want[i], i


def test_png_variants(ctx, ref):
    """Multi-IDAT, ancillary chunks before/after IDAT, palette and RGB images."""
    img = corpus.gradient_noise_rgba(200, 120, 5)
    files = [
        corpus.write_png(img, -1, idat_split=1000),
        corpus.write_png(img, 4, idat_split=7, strategy=zlib.Z_DEFAULT_STRATEGY),
        corpus.write_png(img, 3, extra_chunks=[(b"tEXt", b"Comment\0hello"), (b"gAMA", b"\0\1\x86\xa0")]),
    ]
    res = ctx.decode_png_batch(files)
    for f, (good, w, h, rgba) in zip(files, res):
        rgood, _, _, rrgba = ref.decode_png(f)
        assert good == rgood == 1
        assert rgba == rrgba == img.tobytes()
    # palette (colour type 3): reference is correct here
    r = np.random.default_rng(3)
    pal = r.integers(0, 256, size=(256, 3), dtype=np.uint8)
    idx = (np.add.outer(np.arange(90), np.arange(130)) % 200).astype(np.uint8)[..., None]
    f = corpus.write_png(idx, -1, strategy=zlib.Z_DEFAULT_STRATEGY, palette=pal.tobytes())
    (good, w, h, rgba), = ctx.decode_png_batch([f])
    rgood, _, _, rrgba = ref.decode_png(f)
    assert good == rgood == 1 and rgba == rrgba
    exp = np.concatenate([pal[idx[..., 0]], np.full((90, 130, 1), 255, np.uint8)], axis=2)
    assert rgba == exp.tobytes()
    # RGB (colour type 2): the reference is wrong (D3); this decoder must match the source
    rgb = corpus.gradient_noise_rgba(77, 50, 9)[..., :3].copy()
    f = corpus.write_png(rgb, -1, strategy=zlib.Z_DEFAULT_STRATEGY)
    (good, w, h, rgba), = ctx.decode_png_batch([f])
    exp = np.concatenate([rgb, np.full((50, 77, 1), 255, np.uint8)], axis=2)
    assert good == 1 and rgba == exp.tobytes()


def test_png_unfilter_row_classes_vs_reference(ctx, ref):
    """RGBA8 un-filter by row class (png_unfilter_band4): bands of None/Up rows (columns), None/Sub rows (rows), wavefront
    bands with and without Paeth rows, every mix per band; widths around the 32-pixel block, heights around the 32-row band,
    and images tall enough (> 128 bands) for the hand-off ring between the bands' warps to wrap."""
    rng = np.random.default_rng(11)
    sets = [(0,), (1,), (2,), (3,), (4,), (0, 1), (0, 2), (1, 2), (0, 1, 2, 3), (0, 1, 2, 3, 4)]
    files, want = [], []
    k = 0
    for w, h in ((5, 40), (31, 33), (32, 32), (33, 31), (64, 65), (97, 40), (130, 70), (1024, 96), (40, 5000), (300, 4200)):
        for fs in sets:
            img = corpus.gradient_noise_rgba(w, h, 300 + k, amp=3)  # noisier images trip rule Q12 (compressed > raw)
            if h >= 4000:  # a class per band, so that neighbouring bands of different classes hand rows to each other
                rows = np.repeat([rng.choice(sets[(k + j) % len(sets)]) for j in range((h + 31) // 32)], 32)[:h]
                rows = np.where(rng.random(h) < 0.9, rows, rng.choice(fs, size=h))
            else:
                rows = rng.choice(fs, size=h)
            files.append(corpus.write_png(img, filt=rows, level=1, strategy=zlib.Z_FIXED, single_block=(k % 3 == 0)))
            want.append(img.tobytes())
            k += 1
    res = ctx.decode_png_batch(files)
    for i, (f, (good, w, h, rgba)) in enumerate(zip(files, res)):
        rgood, _, _, rrgba = ref.decode_png(f)
        assert good == rgood, i
        if not good:
            continue
        assert rgba == want[i], i
        if i % 3 == 0:  # single-block streams: the reference's own pixels are right as well (defect D1 needs a late block)
            assert rrgba == want[i], i


def test_empty_and_ragged_batches(ctx):
    assert ctx.inflate_batch([], []) == []
    d = corpus.word_salad(1000, 1)
    s = corpus.raw_deflate(d)
    res = ctx.inflate_batch([s, b"", s[:3], s], [2000, 10, 10, len(s)])  # last: cap == in_size < output
    assert res[0] == (1, d)
    assert res[1][0] == 0 and res[2][0] == 0
    assert res[3][0] == 0  # output overflow fails instead of writing past the buffer


def test_large_batch_roundtrip_property(ctx):
    """Size-independent property at scale: 2048 streams, decode(compress(x)) == x by sha256."""
    base = [corpus.gz_member_cfg2(i, 1 << 16) for i in range(64)]
    members = [base[i % 64] for i in range(2048)]
    res = ctx.decode_gz_batch([g for g, _ in members], [len(d) + len(g) for g, d in members])
    want = [sha(d) for _, d in base]
    for i, (good, out) in enumerate(res):
        assert good == 1 and sha(out) == want[i % 64], i


def test_lane_serial_fixed_block_path_vs_reference(ctx, ref):
    """Large single-fixed-block streams (what stb writes) go through the lane-serial kernels (fx_kernels.cuh):
    one lane per chunk, exact entry points from the 32-hypothesis head pass. BASELINE config 4 shape at reduced size."""
    fx0 = ctx.fx_stats()
    imgs = [corpus.gradient_noise_rgba(1536, 1024, 500 + i) for i in range(3)]
    files = [ref.stb_png(im.tobytes(), 1536, 1024, 4, f) for im, f in zip(imgs, (4, -1, 0))]
    assert all(len(f) > 4 * 32768 for f in files)
    res = ctx.decode_png_batch(files)
    for f, im, (good, w, h, rgba) in zip(files, imgs, res):
        rgood, _, _, rrgba = ref.decode_png(f)
        assert good == rgood
        if good:
            assert rgba == rrgba == im.tobytes()
    # raw inflate of stb zlib streams, incl. long-distance periodic data and a truncated stream
    datas = [corpus.word_salad(900000, 3), corpus.periodic(700000, 4, 30011), bytes(500000)]
    streams = [ref.stb_zlib(d)[2:-4] for d in datas]
    streams.append(streams[0][: len(streams[0]) - 5000])
    caps = [len(d) + len(s) + 64 for d, s in zip(datas + [datas[0]], streams)]
    got = ctx.inflate_batch(streams, caps)
    for s, c, (good, out) in zip(streams, caps, got):
        rgood, rout = ref.inflate(s, c)
        assert good == rgood and out == rout
    fx1 = ctx.fx_stats()
    assert fx1[0] - fx0[0] == 6 and fx1[1] == fx0[1], (fx0, fx1)  # all but the 3 KB stream of zeros took that path, none was handed back


def test_cfg5_extremes_vs_reference(ctx, ref):
    """BASELINE config 5 extremes at full member size (16 MiB): stored random data, period-32000 data (matches
    at the window limit, ratio ~100), long runs and zeros (ratio ~1000), plus small members of every class."""
    big = 16 << 20
    members = [corpus.gz_member_cfg5(i, big) for i in (0, 5, 6, 7)]
    members += [corpus.gz_member_cfg5(i, 65536 + 4099 * i) for i in range(8)]
    caps = [len(d) + len(g) + 64 for g, d in members]
    res = ctx.decode_gz_batch([g for g, _ in members], caps)
    for i, ((g, d), (good, out)) in enumerate(zip(members, res)):
        rgood, rout = ref.decode_gz(g, caps[i])
        assert good == rgood, i
        assert len(out) == len(rout) and sha(out) == sha(rout), i


def test_block_split_long_streams(ctx, ref):
    """Long multi-block streams take the block-split path (chunk-parallel decode behind verified block
    boundaries); results must equal the reference's sequential decode, errors included."""
    rng = np.random.default_rng(77)
    text = corpus.word_salad(5 << 20, 11)
    items = [
        corpus.raw_deflate(text, 6),                                               # ~2 MB, ~70 dynamic blocks
        corpus.raw_deflate(text[: 3 << 20], 1),
        corpus.raw_deflate(corpus.low_entropy(6 << 20, 12), 6),                    # Huffman-only, rule Q2 may end it early
        corpus.raw_deflate(corpus.periodic(6 << 20, 13, 31000), 9),                # distances at the window limit
        corpus.raw_deflate(corpus.runs(40 << 20, 14), 6),                          # ~1000:1
        corpus.mixed_deflate(text[: 4 << 20], 15),                                 # stored + fixed + dynamic segments
        corpus.raw_deflate(bytes(rng.integers(0, 256, 1 << 20, dtype=np.uint8)), 6),  # stored blocks only
        corpus.raw_deflate(text, 6)[:-5000],                                       # truncated inside a block
        corpus.raw_deflate(text[: 200000], 6),                                     # short: stays on the warp-per-stream path
    ]
    broken = bytearray(items[0])
    broken[len(broken) // 2] ^= 0x55                                               # corrupt the middle of a long stream
    items.append(bytes(broken))
    caps = [48 << 20] * len(items)
    before = ctx.bsplit_stats()
    got = ctx.inflate_batch(items, caps)
    after = ctx.bsplit_stats()
    assert after[0] - before[0] >= 5, (before, after)
    for k, (z, (good, out)) in enumerate(zip(items, got)):
        rg, rout = ref.inflate(z, caps[k])
        assert good == rg, k
        if good:
            assert out == rout, k


def test_block_split_false_hints_fall_back(ctx, ref):
    """Stored blocks whose payload is itself DEFLATE data put perfectly valid-looking dynamic block headers
    where no block of the outer stream starts. The chain check must notice and hand the stream back to the
    warp-per-stream kernel; the result is still the reference's."""
    text = corpus.word_salad(3 << 20, 21)
    inner = b"".join(corpus.raw_deflate(corpus.word_salad(8000, 100 + k), 6) for k in range(60))  # ~200 KB of headers
    segs = []
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    segs.append(c.compress(text[: 1 << 20]) + c.flush(zlib.Z_FULL_FLUSH))
    c0 = zlib.compressobj(0, zlib.DEFLATED, -15)
    segs.append(c0.compress(inner) + c0.flush(zlib.Z_FULL_FLUSH))
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    segs.append(c.compress(text[1 << 20:]) + c.flush(zlib.Z_FINISH))
    z = b"".join(segs)
    assert zlib.decompress(z, -15) == text[: 1 << 20] + inner + text[1 << 20:]
    before = ctx.bsplit_stats()
    (good, out), (g2, out2) = ctx.inflate_batch([z, corpus.raw_deflate(text, 6)], [8 << 20, 8 << 20])
    after = ctx.bsplit_stats()
    rg, rout = ref.inflate(z, 8 << 20)
    assert good == rg == 1 and out == rout
    assert g2 == 1 and out2 == text
    assert after[1] - before[1] == 1 and after[0] - before[0] == 1, (before, after)


def test_block_split_equals_warp_per_stream_on_damaged_streams():
    """The block-split path with tiny regions (forced on every stream of >= 64 KiB) against the warp-per-stream
    path of the same library on intact, truncated and bit-flipped streams: same verdict, same bytes. Intact
    streams are also checked against zlib. (The reference itself is not run on damaged input: it has no bounds
    checks there.)"""
    import debigulator_b200 as dbg
    rng = np.random.default_rng(4242)

    def many_blocks(data, level, mem):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, mem)
        return c.compress(data) + c.flush()

    streams, intact = [], []
    for k in range(48):
        n = int(rng.integers(200_000, 1_500_000))
        kind = k % 6
        if kind == 0:
            data = corpus.word_salad(n, 500 + k)
        elif kind == 1:
            data = corpus.low_entropy(n, 500 + k, 5)
        elif kind == 2:
            data = corpus.periodic(n, 500 + k, int(rng.integers(300, 32000)))
        elif kind == 3:
            data = corpus.runs(n * 4, 500 + k)
        elif kind == 4:
            data = corpus.word_salad(n // 2, 500 + k) + bytes(rng.integers(0, 256, n // 8, dtype=np.uint8)) + corpus.word_salad(n // 2, 900 + k)
        else:
            data = corpus.png_filter_rows(corpus.gradient_noise_rgba(512, int(n // 2048) + 8, k), 4)
        z = corpus.mixed_deflate(data, k) if k % 8 == 7 else many_blocks(data, int(rng.choice([1, 6, 9])), int(rng.choice([1, 4, 8, 9])))
        if len(z) < 70_000:
            continue
        streams.append(z)
        intact.append(data)
        if k % 3 == 0:  # truncated
            streams.append(z[: int(len(z) * rng.uniform(0.3, 0.95))])
            intact.append(None)
        if k % 3 == 1:  # a few flipped bits
            b = bytearray(z)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(len(b) // 10, len(b)))] ^= 1 << int(rng.integers(0, 8))
            streams.append(bytes(b))
            intact.append(None)
    caps = [12 << 20] * len(streams)
    for k in range(len(streams)):
        if intact[k] is not None and len(intact[k]) >= len(streams[k]):
            if k % 5 == 0:
                caps[k] = len(intact[k])            # exactly enough room
            elif k % 5 == 1:
                caps[k] = len(intact[k]) - 1 - k    # not enough: both paths must refuse
    old = {k: os.environ.get(k) for k in ("DBG_BSPLIT", "DBG_BSPLIT_REGION", "DBG_BSPLIT_REGION_MIN", "DBG_BSPLIT_MIN_BYTES")}
    try:
        os.environ.update({"DBG_BSPLIT": "1", "DBG_BSPLIT_REGION": "8192", "DBG_BSPLIT_REGION_MIN": "4096", "DBG_BSPLIT_MIN_BYTES": "65536"})
        split_ctx = dbg.Context(0)
        os.environ["DBG_BSPLIT"] = "0"
        plain_ctx = dbg.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    a = split_ctx.inflate_batch(streams, caps)
    b = plain_ctx.inflate_batch(streams, caps)
    taken, fallbacks = split_ctx.bsplit_stats()
    assert plain_ctx.bsplit_stats() == (0, 0) and taken >= len(streams) // 2, (taken, fallbacks)
    for k, ((ga, oa), (gb, ob)) in enumerate(zip(a, b)):
        assert ga == gb, k
        if ga:
            assert oa == ob, k
        if intact[k] is not None:
            if caps[k] < len(intact[k]) and len(intact[k]) >= len(streams[k]) and k % 5 == 1:
                # (a stream that rule Q2 ends early may still fit)
                assert ga == 0 or len(oa) < len(intact[k]), k
            if ga and len(oa) == len(intact[k]):
                assert oa == intact[k], k
    split_ctx.close()
    plain_ctx.close()
