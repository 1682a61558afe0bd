"""Latency of ONE gzip member per compressibility class of corpus.gz_member_cfg5, decoded alone
(device-resident), with and without the block-split path. Usage: time_single_stream.py [MiB]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import debigulator_b200 as dbg  # noqa: E402
from debigulator_b200 import corpus  # noqa: E402

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
ctx = dbg.Context(0)
s = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(s)
i64 = lambda a: torch.from_numpy(np.asarray(a, np.uint64).view(np.int64)).to(dev)
for k in [int(c) for c in os.environ.get("CLASSES", "0,1,2,3,4,5,6,7").split(",")]:
    if k == 8:  # not a config-5 class: stored + fixed + dynamic segments in one stream (config 2's "mixed")
        d = corpus.word_salad(MB << 20, 99)
        g = corpus.gzip_frame(corpus.mixed_deflate(d, 99), d)
    else:
        g, d = corpus.gz_member_cfg5(k, MB << 20)
    h = np.zeros(len(g) + 64, np.uint8)
    h[: len(g)] = np.frombuffer(g, np.uint8)
    cap = len(d) + len(g) + 64
    d_in = torch.from_numpy(h).to(dev)
    d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    sz = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    a = (i64([0]), i64([len(g)]), i64([0]), i64([cap]))
    def step():
        ctx.inflate_device(d_in, a[0], a[1], d_out, a[2], a[3], sz, st, None, stream=s.cuda_stream, gz=True)
    step(); step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"class": k, "compressed": len(g), "out": int(sz.item()), "status": int(st.item()), "ms": round(ms, 3),
                      "MBps_out": round(int(sz.item()) / ms / 1e3, 1), "bsplit": os.environ.get("DBG_BSPLIT", "1"),
                      "bsplit_stats": ctx.bsplit_stats()}))
