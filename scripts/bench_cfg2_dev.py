"""cfg2 device-resident timing only (for ncu launch lists). usage: bench_cfg2_dev.py [members]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import debigulator_b200 as dbg
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
uniq = bench.make_unique(bench._gen_gz, 64)
classes = [int(c) for c in os.environ.get("CLASSES", "0,1,2,3").split(",")]
uniq = [u for i, u in enumerate(uniq) if i % 4 in classes]
offs, sizes, total = bench.pack([u[0] for u in uniq], n)
h = np.zeros(total + 64, np.uint8)
for i in range(n):
    b = uniq[i % len(uniq)][0]; h[offs[i]:offs[i] + len(b)] = np.frombuffer(b, np.uint8)
dev = torch.device("cuda", 0); ctx = dbg.Context(0)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
stride = (1 << 20) + 4096
i64 = lambda a: torch.from_numpy(np.asarray(a, np.uint64).view(np.int64)).to(dev)
d_in = torch.from_numpy(h).to(dev); d_out = torch.zeros(n * stride, dtype=torch.uint8, device=dev)
a_off, a_sz, o_off, o_cap = i64(offs), i64(sizes), i64(np.arange(n, dtype=np.uint64) * np.uint64(stride)), i64(np.full(n, stride, np.uint64))
osz = torch.zeros(n, dtype=torch.int64, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
def step(): ctx.inflate_device(d_in, a_off, a_sz, d_out, o_off, o_cap, osz, st, None, stream=s.cuda_stream, gz=True)
step(); step(); torch.cuda.synchronize()
assert int(st.abs().sum()) == 0 and bool((osz == (1 << 20)).all())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(json.dumps({"members": n, "ms": ms, "GBps": n * (1 << 20) / ms / 1e6, "classes": classes, "bsplit": ctx.bsplit_stats(), "fx": ctx.fx_stats(), "lanes": ctx.lane_stats()}))
