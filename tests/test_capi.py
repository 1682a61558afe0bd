"""CPU suite, part 3: the C-ABI library builds for sm_100a, loads, exports every
symbol the headers declare, and refuses to work without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

import debigulator_b200 as dbg
from debigulator_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def declared_symbols():
    names = set()
    for h in os.listdir(INC):
        src = open(os.path.join(INC, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src):
            n = m.group(1)
            if n.startswith(("dbg_", "inflate", "decode_", "init_", "get_PNG")):
                names.add(n)
    return names


def test_library_builds_and_exports_all_declared_symbols():
    lib = build_library()
    assert os.path.exists(lib)
    L = C.CDLL(lib, mode=C.RTLD_LOCAL)
    want = declared_symbols()
    assert {"inflate", "inflate_init", "decode_png", "decode_png_init", "decode_png_get_width_height", "decode_gz",
            "init_decode_gz", "decode_PNG", "dbg_inflate_batch", "dbg_decode_png_batch", "dbg_decode_gz_batch",
            "dbg_decode_batch_packed", "dbg_inflate_batch_device"} <= want
    missing = [n for n in sorted(want) if not hasattr(L, n)]
    assert not missing, missing


def test_get_width_height_host_only(golden_dir):
    data = open(os.path.join(golden_dir, "gimp_test.png"), "rb").read()
    assert dbg.png_get_width_height(data) == (1, 1024, 1024)
    assert dbg.png_get_width_height(data[:20])[0] == 0          # decode_png.c:627
    assert dbg.png_get_width_height(b"\x89QNG" + data[4:])[0] == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = dbg.load_library()
    assert L.dbg_device_count() < 0
    with pytest.raises(dbg.DebigulatorError):
        dbg.Context(0)
    # batch entry points refuse a NULL context instead of decoding on the host
    assert L.dbg_inflate_batch(None, 1, None, None, None, None, None, None) == -1


def test_sass_is_sm100a_only():
    import subprocess
    lib = build_library()
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_corpus_single_block_encoder_roundtrip():
    """The corpus tool that writes stb-shaped streams (one final fixed-Huffman block) is a valid DEFLATE encoder."""
    import zlib
    from debigulator_b200 import corpus
    for seed, n in ((1, 0), (2, 1), (3, 70000), (4, 300000)):
        data = corpus.word_salad(n, seed) if n else b""
        z = corpus.fixed_block_deflate(data)
        assert (z[0] & 7) == 3                      # BFINAL=1, BTYPE=01
        assert zlib.decompress(z, -15) == data
    img = corpus.gradient_noise_rgba(64, 48, 9)
    png = corpus.write_png(img, 4, single_block=True)
    from PIL import Image
    import io
    assert Image.open(io.BytesIO(png)).convert("RGBA").tobytes() == img.tobytes()
