"""BASELINE config 1 on the GPU path: hellopng-style decode of tests/golden/gimp_test.png and hellogz-style decode
of gzipsample.gz through the host batch API (batch of one), wall clock incl. H2D/D2H, beside the reference C."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import debigulator_b200 as dbg  # noqa: E402
from oracle import checker  # noqa: E402

png = open(os.path.join(ROOT, "tests/golden/gimp_test.png"), "rb").read()
gz = open(os.path.join(ROOT, "tests/golden/gzipsample.gz"), "rb").read()
ctx = dbg.Context(0)
out = {}
for name, fn, ref in (("gimp_test.png", lambda: ctx.decode_png_batch([png]), lambda: checker.decode_png(png)),
                      ("gzipsample.gz", lambda: ctx.decode_gz_batch([gz], [700000]), lambda: checker.decode_gz(gz, 700000))):
    fn()
    t = []
    for _ in range(10):
        t0 = time.perf_counter()
        r = fn()
        t.append(time.perf_counter() - t0)
    assert r[0][0] == 1
    ref()
    t0 = time.perf_counter()
    for _ in range(5):
        ref()
    tr = (time.perf_counter() - t0) / 5
    out[name] = {"gpu_ms_median": sorted(t)[5] * 1e3, "gpu_ms_best": min(t) * 1e3, "reference_c_ms": tr * 1e3}
print(json.dumps(out))
