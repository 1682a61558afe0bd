"""Wave timelines (DBG_WAVE_TRACE) of cfg2 through the pipe: two sub-batches in flight. usage: trace_pipe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
import debigulator_b200 as dbg
import bench
n = 4096
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
dt, osz, st, l = bench.e2e_pipelined(dbg, 0, dbg.api.KIND_GZ, hi, a[0], a[1], ho, a[2], a[3], 1, lambda: None, lambda x: x, parts=2, depth=2)
os.environ["DBG_WAVE_TRACE"] = "1"
dt, osz, st, l = bench.e2e_pipelined(dbg, 0, dbg.api.KIND_GZ, hi, a[0], a[1], ho, a[2], a[3], 2, lambda: None, lambda x: x, parts=2, depth=2)
print("ms per step", dt * 1e3)
