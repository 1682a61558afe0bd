"""BASELINE config 5 shape, scaled: gzip members with sizes log-uniform in [64 KiB, 16 MiB] and the
compressibility sweep of corpus.gz_member_cfg5 (stored / text / low entropy / period ~32 kB / runs / zeros),
decoded device-resident. Reports decompressed GB/s. Usage: bench_cfg5.py [n_members] [n_unique]"""
import json
import multiprocessing as mp
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import debigulator_b200 as dbg  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
U = int(sys.argv[2]) if len(sys.argv) > 2 else 64


def gen(i):
    from debigulator_b200 import corpus
    size = int(65536 * (256.0 ** (((i * 2654435761) % 1000) / 999.0)))
    g, d = corpus.gz_member_cfg5(i, size)
    return g, len(d), zlib.crc32(d)


with mp.get_context("fork").Pool(min(U, len(os.sched_getaffinity(0)))) as pool:
    uniq = pool.map(gen, range(U))
dev = torch.device("cuda", 0)
ctx = dbg.Context(0)
offs, sizes, caps, total = [], [], [], 0
for i in range(N):
    g, n, _ = uniq[i % U]
    offs.append(total)
    sizes.append(len(g))
    caps.append((n + len(g) + 64 + 15) // 16 * 16)
    total += (len(g) + 31) // 16 * 16
h = np.zeros(total + 64, np.uint8)
for i in range(N):
    g = uniq[i % U][0]
    h[offs[i]:offs[i] + len(g)] = np.frombuffer(g, np.uint8)
out_off = np.concatenate([[0], np.cumsum(caps[:-1])]).astype(np.uint64)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
i64 = lambda a: torch.from_numpy(np.asarray(a, np.uint64).view(np.int64)).to(dev)
d_in = torch.from_numpy(h).to(dev)
d_out = torch.zeros(int(sum(caps)), dtype=torch.uint8, device=dev)
d_size = torch.zeros(N, dtype=torch.int64, device=dev)
d_st = torch.zeros(N, dtype=torch.int32, device=dev)
order = np.argsort(-np.asarray(sizes, np.int64), kind="stable").astype(np.uint32)
d_order = torch.from_numpy(order.view(np.int32)).to(dev) if os.environ.get("CFG5_HOST_ORDER") else None  # None: the library orders the queue
a_off, a_sz, o_off, o_cap = i64(offs), i64(sizes), i64(out_off), i64(caps)


def step():
    ctx.inflate_device(d_in, a_off, a_sz, d_out, o_off, o_cap, d_size, d_st, d_order, stream=s.cuda_stream, gz=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
assert int(d_st.abs().sum()) == 0
got = d_size.cpu().numpy()
exact = sum(int(got[i]) == uniq[i % U][1] for i in range(N))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
out_bytes = int(got.sum())
print(json.dumps({"workload": f"cfg5 shape: {N} gzip members, 64 KiB-16 MiB log-uniform, 8 compressibility classes, {U} unique",
                  "output_bytes": out_bytes, "compressed_bytes": int(sum(sizes)), "ms_per_step": ms,
                  "GBps_out": out_bytes / ms / 1e6, "members_with_spec_size": exact, "members": N,
                  "bsplit": os.environ.get("DBG_BSPLIT", "1"), "bsplit_streams_fallbacks": ctx.bsplit_stats()}))
