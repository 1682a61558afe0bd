"""Wave timeline of the packed host API (DBG_WAVE_TRACE=1) for a PNG or gzip batch, and batch-of-one latencies.
usage: trace_e2e.py png|gz|one"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DBG_WAVE_TRACE"] = "1"
import numpy as np, torch
import debigulator_b200 as dbg
from debigulator_b200 import corpus
import bench
what = sys.argv[1]
ctx = dbg.Context(0)
if what == "one":
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    gz = open(os.path.join(gold, "gzipsample.gz"), "rb").read()
    png = open(os.path.join(gold, "gimp_test.png"), "rb").read()
    for name, fn in (("gz", lambda: ctx.decode_gz_batch([gz], [600000])), ("png", lambda: ctx.decode_png_batch([png]))):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
        print(name, "ms:", " ".join("%.2f" % t for t in ts))
    sys.exit(0)
if what == "png":
    n, w, h = 2048, 1024, 1024
    uniq = bench.pool_map(bench._gen_png, [(i, w, h, None) for i in range(12)])
    kind, caps = dbg.api.KIND_PNG, [w * h * 4] * n
else:
    n = 4096
    uniq = bench.make_unique(bench._gen_gz, 64)
    kind, caps = dbg.api.KIND_GZ, [(1 << 20) + 4096] * n
offs, sizes, total = bench.pack([u[0] for u in uniq], n)
h_in = torch.empty(total + 64, dtype=torch.uint8).pin_memory(); hi = h_in.numpy(); hi[:] = 0
for i in range(n):
    b = uniq[i % len(uniq)][0]; hi[offs[i]:offs[i] + len(b)] = np.frombuffer(b, np.uint8)
out_off = np.concatenate([[0], np.cumsum(caps[:-1])]).astype(np.uint64)
h_out = torch.empty(int(sum(caps)), dtype=torch.uint8).pin_memory(); ho = h_out.numpy()
a = (np.asarray(offs, np.uint64), np.asarray(sizes, np.uint64), out_off, np.asarray(caps, np.uint64))
for rep in range(3):
    t0 = time.perf_counter()
    osz, st = ctx.decode_packed(kind, hi, a[0], a[1], ho, a[2], a[3])
    print("call %d: %.1f ms, failures %d" % (rep, (time.perf_counter() - t0) * 1e3, int((st != 0).sum())), file=sys.stderr)
