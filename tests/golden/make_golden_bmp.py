"""Regenerates tests/golden/bmp_manifest.json by running the UNMODIFIED reference decode_bmp.c
(oracle/_ref/libref.so, built in place from /root/reference by oracle/Makefile) on
  * the three BMP fixtures copied from /root/reference/resources (data, not code),
  * small synthetic BMPs (both row orders, 40- and 108-byte DIB headers, shifted pixel offsets),
  * header mutations the reference rejects (decode_bmp.c:120-221),
and its encode_BMP on the decoded fixtures. Run in the build container only:

    python tests/golden/make_golden_bmp.py
"""
import base64
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reflib  # noqa: E402
from debigulator_b200 import corpus  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sha = lambda b: hashlib.sha256(b).hexdigest()


def entry(data, inline=True):
    g, w, h, rgba = reflib.decode_bmp(data)
    gd, dw, dh = reflib.bmp_dims(data)
    e = {"good": g, "w": w, "h": h, "dims_good": gd, "rgba_sha256": sha(rgba) if g else None, "bytes": len(data)}
    if g:
        n, enc = reflib.encode_bmp(rgba, w, h)
        e["encode_size"] = n
        e["encode_sha256"] = sha(enc)
    if inline:
        e["b64"] = base64.b64encode(data).decode()
    return e


def main():
    out = {"generator": "tests/golden/make_golden_bmp.py (reference = oracle/_ref/libref.so, silent no-assert build)",
           "fixtures": {}, "synthetic": {}}
    for name in ("structuredart.bmp", "fs_psychologist.bmp", "fs_fightingpit.bmp"):
        out["fixtures"][name] = entry(open(os.path.join(HERE, name), "rb").read(), inline=False)
    rng = np.random.default_rng(0xB3B)
    def img(w, h):
        return rng.integers(0, 256, w * h * 4, dtype=np.uint8).tobytes()
    cases = {
        "top_down_7x5": corpus.bmp_file(img(7, 5), 7, 5),
        "bottom_up_7x5": corpus.bmp_file(img(7, 5), 7, 5, bottom_up=True),
        "v4_bottom_up_9x4": corpus.bmp_file(img(9, 4), 9, 4, bottom_up=True, v4=True),
        "odd_offset_6x6": corpus.bmp_file(img(6, 6), 6, 6, bottom_up=True, pad=1),
        "aligned_offset_33x3": corpus.bmp_file(img(33, 3), 33, 3, pad=2),
        "one_pixel": corpus.bmp_file(img(1, 1), 1, 1, bottom_up=True),
        "empty_0x0": corpus.bmp_file(b"", 0, 0),
        "bad_magic": corpus.bmp_file(img(4, 4), 4, 4, magic=b"BA"),
        "dib_124": corpus.bmp_file(img(4, 4), 4, 4, dib_size=124),
        "planes_2": corpus.bmp_file(img(4, 4), 4, 4, planes=2),
        "bpp_24": corpus.bmp_file(img(4, 4), 4, 4, bpp=24),
        "file_longer_than_declared": corpus.bmp_file(img(4, 4), 4, 4, bf_size=10) + bytes(80),
        "file_shorter_than_declared": corpus.bmp_file(img(4, 4), 4, 4, bf_size=100000),
    }
    for k, v in cases.items():
        out["synthetic"][k] = entry(v)
    with open(os.path.join(HERE, "bmp_manifest.json"), "w") as f:
        json.dump(out, f, indent=1)
    for grp in ("fixtures", "synthetic"):
        for k, v in out[grp].items():
            print(grp, k, v["good"], v["w"], v["h"], v["dims_good"])


if __name__ == "__main__":
    main()
