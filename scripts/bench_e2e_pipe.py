"""cfg2 end to end: one blocking packed call per step against the pipe (dbg_pipe_*), a few depths / splits.
usage: bench_e2e_pipe.py [members]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
import debigulator_b200 as dbg
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
uniq = bench.make_unique(bench._gen_gz, 64)
size = 1 << 20
offs, sizes, total = bench.pack([u[0] for u in uniq], n)
h_in = torch.empty(total + 64, dtype=torch.uint8).pin_memory(); hi = h_in.numpy(); hi[:] = 0
for i in range(n):
    b = uniq[i % 64][0]; hi[offs[i]:offs[i] + len(b)] = np.frombuffer(b, np.uint8)
stride = size + 4096
out_off = np.arange(n, dtype=np.uint64) * np.uint64(stride)
caps = np.full(n, stride, np.uint64)
h_out = torch.empty(n * stride, dtype=torch.uint8).pin_memory(); ho = h_out.numpy()
a = (np.asarray(offs, np.uint64), np.asarray(sizes, np.uint64), out_off, caps)
ctx = dbg.Context(0)
ctx.decode_packed(dbg.api.KIND_GZ, hi, *a[:2], ho, *a[2:])
t0 = time.perf_counter()
for _ in range(3): ctx.decode_packed(dbg.api.KIND_GZ, hi, *a[:2], ho, *a[2:])
print("blocking call: %.1f ms per step" % ((time.perf_counter() - t0) / 3 * 1e3))
ctx.trim()
for depth, parts in ((2, 2), (2, 4), (3, 3), (4, 4)):
    dt, osz, st, l = bench.e2e_pipelined(dbg, 0, dbg.api.KIND_GZ, hi, a[0], a[1], ho, a[2], a[3], 4, lambda: None, lambda x: x, parts=parts, depth=depth)
    ok = int(st.sum()) == 0
    print("pipe depth %d, %d sub-batches per step: %.1f ms per step (%.1f GB/s) ok=%s" % (depth, parts, dt * 1e3, n * size / dt / 1e9, ok))
