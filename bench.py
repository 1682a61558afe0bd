#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched inflate / gzip / PNG decode path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input. The
headline workload is BASELINE.json config 2: a batch of 4096 synthetic 1 MiB
gzip members per GPU (classes stored / fixed / dynamic / mixed by i % 4,
SURVEY.md 8d), decoded to their payloads. Weak scaling: every rank decodes its
own 4096-member batch; no collective is needed on the data path.

  value   decompressed GB/s, whole job, inputs resident in HBM, timed with CUDA
          events on the launching stream, max over ranks
  e2e     the same metric through the public C-ABI with pinned HOST buffers:
          H2D of the compressed batch + kernels + D2H of the payloads per step
  roofline / cpu_baseline / clocks / png : see DESIGN.md "Measurement"

`--impl reference` times the reference's own CPU implementation (oracle/_ref,
the unmodified C sources; else the oracle port) on all host cores, on a bounded
sample of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MEMBER_BYTES = 1 << 20
N_MEMBERS = 4096
N_UNIQUE = 64
PNG_N, PNG_W, PNG_H, PNG_UNIQUE = 2048, 1024, 1024, 12
METRIC = "inflate_output_GBps"
UNIT = "GB/s"


def _gen_gz(i):
    from debigulator_b200 import corpus
    return corpus.gz_member_cfg2(i, MEMBER_BYTES)


def _gen_png(i):
    from debigulator_b200 import corpus
    return corpus.png_cfg3(i, PNG_W, PNG_H)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_unique(gen, count):
    workers = max(1, min(host_cores(), count))
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(gen, range(count))


def pack(items, n, align=16, pad=16):
    """Cycles `items` (compressed bytes) to n entries inside one arena."""
    offs, sizes, total = [], [], 0
    for i in range(n):
        b = items[i % len(items)]
        offs.append(total)
        sizes.append(len(b))
        total += (len(b) + pad + align - 1) // align * align
    return offs, sizes, total


# ---------------------------------------------------------------- CPU baseline --
def _cpu_worker(args):
    kind, blobs, caps = args
    from oracle import checker
    t0 = time.perf_counter()
    out_bytes = 0
    for b, c in zip(blobs, caps):
        if kind == "gz":
            good, out = checker.decode_gz(b, c)
        else:
            good, _, _, out = checker.decode_png(b)
        assert good == 1
        out_bytes += len(out)
    return out_bytes, time.perf_counter() - t0


def cpu_throughput(kind, blobs, caps, procs):
    """One process per core (each with its own reference slot 0), static partition."""
    shards = [([], []) for _ in range(procs)]
    for i, (b, c) in enumerate(zip(blobs, caps)):
        shards[i % procs][0].append(b)
        shards[i % procs][1].append(c)
    shards = [s for s in shards if s[0]]
    t0 = time.perf_counter()
    if len(shards) == 1:
        res = [_cpu_worker((kind, shards[0][0], shards[0][1]))]
    else:
        with mp.get_context("fork").Pool(len(shards)) as pool:
            res = pool.map(_cpu_worker, [(kind, s[0], s[1]) for s in shards])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


# ------------------------------------------------------------------- clocks -----
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ reference ---
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import checker
    cores = host_cores()
    uniq = make_unique(_gen_gz, N_UNIQUE)
    per_step = max(cores * 8, 64)  # members per step: a bounded sample of the 4096-member batch
    blobs = [uniq[i % N_UNIQUE][0] for i in range(per_step)]
    caps = [MEMBER_BYTES + len(b) for b in blobs]
    for _ in range(args.warmup):
        cpu_throughput("gz", blobs[:cores], caps[:cores], cores)
    tot_bytes, tot_wall = 0, 0.0
    for _ in range(args.steps):
        b, w = cpu_throughput("gz", blobs, caps, cores)
        tot_bytes += b
        tot_wall += w
    val = tot_bytes / tot_wall / 1e9
    sample = f"{per_step} of {N_MEMBERS} members per step ({per_step * MEMBER_BYTES >> 20} MiB out), one process per core"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "cfg2: 4096 x 1 MiB gzip members (stored/fixed/dynamic/mixed)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": checker.kind(), "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------- ours ---
def run_ours(args):
    import torch
    import torch.distributed as dist
    import debigulator_b200 as dbg
    from debigulator_b200.build import build_library

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    build_library()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = dbg.Context(local)
    n, size = args.members, MEMBER_BYTES
    if args.png_only:
        tstream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(tstream)
        print(json.dumps({"png": bench_png(ctx, dev, torch, args.png_images)}))
        return

    # ---- corpus: N_UNIQUE distinct members (16 per class), cycled to n, every copy at its own address
    uniq = make_unique(_gen_gz, N_UNIQUE)
    offs, sizes, in_total = pack([u[0] for u in uniq], n)
    h_in = torch.empty(in_total + 64, dtype=torch.uint8).pin_memory()
    h_in.zero_()
    hv = h_in.numpy()
    for i in range(n):
        b = uniq[i % N_UNIQUE][0]
        hv[offs[i]:offs[i] + len(b)] = np.frombuffer(b, np.uint8)
    # inflate() requires recipient_size >= compressed_input_size (inflate.c:826); stored members are
    # slightly larger than their payload, so every output slot carries 4 KiB of slack
    stride = size + 4096
    out_total = n * size
    out_span = n * stride
    h_out = torch.empty(out_span, dtype=torch.uint8).pin_memory()
    in_off = np.array(offs, dtype=np.uint64)
    in_size = np.array(sizes, dtype=np.uint64)
    out_off = (np.arange(n, dtype=np.uint64) * np.uint64(stride))
    out_cap = np.full(n, stride, dtype=np.uint64)
    comp_bytes = int(in_size.sum())

    tstream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: kernels, events and checks all live on it
    torch.cuda.set_stream(tstream)
    d_in = h_in.to(dev, non_blocking=False)
    d_out = torch.zeros(out_span, dtype=torch.uint8, device=dev)
    t_i64 = lambda a: torch.from_numpy(a.view(np.int64)).to(dev)
    d_in_off, d_in_size, d_out_off, d_out_cap = t_i64(in_off), t_i64(in_size), t_i64(out_off), t_i64(out_cap)
    d_out_size = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    # scheduling hint (any caller can compute it from the first deflate byte): longest-processing-time first,
    # where a member that starts with a stored block counts as 1/16 of its size
    first = np.array([uniq[i % N_UNIQUE][0][10] & 6 for i in range(n)])
    weight = np.where(first == 0, in_size // 16, in_size).astype(np.int64)
    order = np.argsort(-weight, kind="stable").astype(np.uint32)
    d_order = torch.from_numpy(order.view(np.int32)).to(dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        ctx.inflate_device(d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_out_size, d_status, d_order,
                           stream=stream, gz=True)

    def verify():
        torch.cuda.synchronize()
        bad = torch.nonzero(d_status).flatten()
        assert bad.numel() == 0, f"device decode reported failures: {[(int(i), int(d_status[i])) for i in bad[:8]]}"
        assert bool((d_out_size == size).all().item()), "wrong output sizes"
        exp = torch.stack([torch.from_numpy(np.frombuffer(u[1], np.uint8).copy()) for u in uniq]).to(dev)
        got = d_out.view(n, stride)
        idx = torch.arange(n, device=dev) % N_UNIQUE
        for s in range(0, n, 256):
            assert torch.equal(got[s:s + 256, :size], exp[idx[s:s + 256]]), "payload mismatch"

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_device()
    barrier()
    verify()
    d_out.zero_()
    launches0 = ctx.kernel_launches
    ctx.profile_enable(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    kern_ms, kern_n = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.kernel_launches - launches0
    verify()
    value = world * n * size * args.steps / (dev_ms / 1e3) / 1e9

    # ---- end to end through the packed host API (pinned host arenas)
    hin_np, hout_np = h_in.numpy(), h_out.numpy()
    e2e_steps = max(1, min(args.steps, 5))
    ctx.decode_packed(dbg.api.KIND_GZ, hin_np, in_off, in_size, hout_np, out_off, out_cap)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        osz, st = ctx.decode_packed(dbg.api.KIND_GZ, hin_np, in_off, in_size, hout_np, out_off, out_cap)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert int(st.sum()) == 0 and int(osz.sum()) == out_total
    for k in (0, 1, 2, 3, n - 1):
        assert hout_np[k * stride:k * stride + size].tobytes() == uniq[k % N_UNIQUE][1], "e2e payload mismatch"
    e2e_val = world * out_total * e2e_steps / e2e_s / 1e9
    # the link ceiling of that path: the same pinned arenas copied H2D and D2H at once, nothing else running
    pcie = None
    if rank == 0 and world == 1:
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2
        pcie = {"h2d_plus_d2h_s": dt, "output_GBps_if_copies_only": out_total / dt / 1e9}

    # ---- PNG (BASELINE config 3 shape), secondary metric
    png = None
    if args.png and rank == 0 and world == 1:
        png = bench_png(ctx, dev, torch, args.png_images)

    # ---- PNG, BASELINE config 4 shape (few huge images): exercises the split-stream path
    png_large = None
    if args.png_large and rank == 0 and world == 1:
        png_large = bench_png(ctx, dev, torch, args.png_large, 8192, 8192, 1)
    bmp = None
    if args.bmp and rank == 0 and world == 1:
        bmp = bench_bmp(ctx, dev, torch, args.bmp)
    cfg5 = None
    if args.cfg5 and rank == 0 and world == 1:
        cfg5 = bench_cfg5(ctx, dev, torch, args.cfg5)

    # ---- CPU baseline: the reference C on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import checker
        cores = host_cores()
        per = max(cores * 8, 64)
        blobs = [uniq[i % N_UNIQUE][0] for i in range(per)]
        caps = [size + len(b) for b in blobs]
        b1, w1 = cpu_throughput("gz", blobs[:16], caps[:16], 1)
        bn, wn = cpu_throughput("gz", blobs, caps, cores)
        cpu = {"value": bn / wn / 1e9, "unit": UNIT, "cores": cores, "kind": checker.kind(),
               "sample": f"{per} of {n} members ({per} MiB out), one process per core",
               "single_thread": {"value": b1 / w1 / 1e9, "unit": UNIT, "sample": "16 members (4 per class)"}}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = comp_bytes + out_total  # C + U per launch
        traffic = None  # dram read + write of one inflate launch from the committed `ncu --set full` capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_inflate_traffic.json")))
            if n == N_MEMBERS and int(tj["algorithmic_bytes_per_launch"]) == int(alg_bytes):
                traffic = float(tj["traffic_bytes_per_launch"])
        except Exception:
            pass
        avg_kern_ms = kern_ms / max(kern_n, 1)
        achieved = alg_bytes / (avg_kern_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"cfg2: {n} x 1 MiB gzip members per GPU (stored/fixed/dynamic/mixed by i%4)",
                       "unique_members": N_UNIQUE, "compressed_bytes_per_gpu": comp_bytes, "output_bytes_per_gpu": out_total,
                       "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no explicit flush" % ((comp_bytes + out_total) / 1e9),
                       "parallelism": f"{world} independent shard(s), no collective"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(in_total + 64), "d2h_bytes_per_step": int(out_span),
                    "steps": e2e_steps, "api": "dbg_decode_batch_packed(kind=gzip), pinned host arenas",
                    "link_ceiling": pcie},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles/r01_inflate_traffic.json (ncu --set full)" if traffic else None,
                         "kernel": "inflate_batch_kernel", "avg_launch_ms": avg_kern_ms, "launches": kern_n,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        if png:
            line["png"] = png
        if png_large:
            line["png_cfg4_shape"] = png_large
        if cfg5:
            line["cfg5_shape"] = cfg5
        if bmp:
            bmp["roofline"]["peak"] = peak
            bmp["roofline"]["frac"] = bmp["roofline"]["achieved"] / peak
            line["bmp"] = bmp
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _gen_png_large(i):
    from debigulator_b200 import corpus
    return corpus.png_cfg3(6 * i + 5, 8192, 8192)  # forced Paeth, as BASELINE config 4


def bench_png(ctx, dev, torch, n=PNG_N, w=PNG_W, h=PNG_H, n_unique=PNG_UNIQUE):
    PNG_W_, PNG_H_ = w, h
    uniq = make_unique(_gen_png if (w, h) == (PNG_W, PNG_H) else _gen_png_large, n_unique)
    PNG_UNIQUE_ = n_unique
    offs, sizes, in_total = pack([u[0] for u in uniq], n)
    h_in = np.zeros(in_total + 64, dtype=np.uint8)
    for i in range(n):
        b = uniq[i % PNG_UNIQUE_][0]
        h_in[offs[i]:offs[i] + len(b)] = np.frombuffer(b, np.uint8)
    rgba = PNG_W_ * PNG_H_ * 4
    d_in = torch.from_numpy(h_in).to(dev)
    d_out = torch.zeros(n * rgba, dtype=torch.uint8, device=dev)
    i64 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.uint64).view(np.int64)).to(dev)
    d_in_off, d_in_size = i64(offs), i64(sizes)
    d_out_off, d_out_cap = i64(np.arange(n, dtype=np.uint64) * np.uint64(rgba)), i64(np.full(n, rgba, dtype=np.uint64))
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    tot_in = int(sum(sizes))

    def step():
        ctx.png_device(d_in, d_in_off, d_in_size, d_out, d_out_off, d_out_cap, d_status, tot_in, n * rgba, stream=stream)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    assert int(d_status.abs().sum().item()) == 0, "png decode failures"
    exp = torch.stack([torch.from_numpy(np.frombuffer(u[1], np.uint8).copy()) for u in uniq]).to(dev)
    got = d_out.view(n, rgba)
    idx = torch.arange(n, device=dev) % PNG_UNIQUE_
    step_chk = max(1, (1 << 28) // rgba)
    for s in range(0, n, step_chk):
        assert torch.equal(got[s:s + step_chk], exp[idx[s:s + step_chk]]), "png pixel mismatch"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    shape = "cfg3 shape" if (w, h) == (PNG_W, PNG_H) else "cfg4 shape (forced Paeth; the config asks for 32 per GPU)"
    filt = "filters None/Sub/Up/Avg/Paeth/adaptive by i%6, " if (w, h) == (PNG_W, PNG_H) else ""
    return {"metric": "png_decode_Mpixels_per_s", "value": n * PNG_W_ * PNG_H_ / (ms / 1e3) / 1e6, "unit": "Mpix/s",
            "rgba_GBps": n * rgba / (ms / 1e3) / 1e9, "ms_per_step": ms,
            "config": {"workload": f"{shape}: {n} x {PNG_W_}x{PNG_H_} RGBA PNGs, {filt}one fixed-Huffman block each (stb stream shape)",
                       "unique_images": PNG_UNIQUE_, "png_bytes": tot_in}}


def _gen_cfg5(i):
    from debigulator_b200 import corpus
    size = int(65536 * (256.0 ** (((i * 2654435761) % 1000) / 999.0)))  # log-uniform in [64 KiB, 16 MiB]
    g, d = corpus.gz_member_cfg5(i, size)
    return g, len(d)


def bench_cfg5(ctx, dev, torch, n, n_unique=64):
    """BASELINE config 5 shape, scaled to one GPU: gzip members of 64 KiB-16 MiB (log-uniform), eight
    compressibility classes, device-resident. The long members take the block-split path."""
    with mp.get_context("fork").Pool(min(n_unique, host_cores())) as pool:
        uniq = pool.map(_gen_cfg5, range(n_unique))
    offs, sizes, caps, total = [], [], [], 0
    for i in range(n):
        g, m = uniq[i % n_unique]
        offs.append(total)
        sizes.append(len(g))
        caps.append((m + len(g) + 64 + 15) // 16 * 16)
        total += (len(g) + 31) // 16 * 16
    h = np.zeros(total + 64, np.uint8)
    for i in range(n):
        g = uniq[i % n_unique][0]
        h[offs[i]:offs[i] + len(g)] = np.frombuffer(g, np.uint8)
    out_off = np.concatenate([[0], np.cumsum(caps[:-1])]).astype(np.uint64)
    i64 = lambda a: torch.from_numpy(np.asarray(a, np.uint64).view(np.int64)).to(dev)
    d_in = torch.from_numpy(h).to(dev)
    d_out = torch.zeros(int(sum(caps)), dtype=torch.uint8, device=dev)
    d_size = torch.zeros(n, dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    a_off, a_sz, o_off, o_cap = i64(offs), i64(sizes), i64(out_off), i64(caps)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.inflate_device(d_in, a_off, a_sz, d_out, o_off, o_cap, d_size, d_st, None, stream=stream, gz=True)

    before = ctx.bsplit_stats()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    assert int(d_st.abs().sum().item()) == 0, "cfg5 decode failures"
    got = d_size.cpu().numpy()
    # spot check of the longest text member against zlib
    k = max(range(min(n, n_unique)), key=lambda i: sizes[i] if i % 8 in (1, 2) else 0)
    want = zlib.decompress(uniq[k][0], 31)[: int(got[k])]
    assert d_out[int(out_off[k]):int(out_off[k]) + int(got[k])].cpu().numpy().tobytes() == want, "cfg5 payload mismatch"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    after = ctx.bsplit_stats()
    out_bytes = int(got.sum())
    return {"metric": "inflate_output_GBps", "value": out_bytes / ms / 1e6, "unit": "GB/s", "ms_per_step": ms,
            "config": {"workload": f"cfg5 shape: {n} gzip members, 64 KiB-16 MiB log-uniform, 8 compressibility classes "
                                   "(stored / text / Huffman-only / period ~32 kB / runs / zeros)",
                       "unique_members": n_unique, "output_bytes": out_bytes, "compressed_bytes": int(sum(sizes)),
                       "members_with_spec_size": int(sum(int(got[i]) == uniq[i % n_unique][1] for i in range(n)))},
            "block_split_streams_per_step": (after[0] - before[0]) // (steps + 2), "block_split_fallbacks": after[1] - before[1]}


def bench_bmp(ctx, dev, torch, n, w=2048, h=2048):
    """decode_BMP / encode_BMP on n bottom-up w x h files: the one kernel of the
    library that is plain HBM traffic (every pixel read once, written once)."""
    from debigulator_b200 import corpus
    rng = np.random.default_rng(7)
    rgba = rng.integers(0, 256, w * h * 4, dtype=np.uint8).tobytes()
    file = corpus.bmp_file(rgba, w, h, bottom_up=True)
    fsz, osz = len(file), w * h * 4
    stride = (fsz + 255) // 256 * 256
    one = torch.from_numpy(np.frombuffer(file + bytes(stride - fsz), np.uint8).copy()).to(dev)
    d_in = one.repeat(n)
    d_out = torch.zeros(n * osz, dtype=torch.uint8, device=dev)
    i64 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.uint64).view(np.int64)).to(dev)
    in_off, in_size = i64(np.arange(n, dtype=np.uint64) * np.uint64(stride)), i64(np.full(n, fsz, dtype=np.uint64))
    out_off, out_cap = i64(np.arange(n, dtype=np.uint64) * np.uint64(osz)), i64(np.full(n, osz, dtype=np.uint64))
    out_size = torch.zeros(n, dtype=torch.int64, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    # encode target: the decoded images back into BMP files
    estride = (54 + osz + 1 + 255) // 256 * 256
    d_enc = torch.zeros(n * estride, dtype=torch.uint8, device=dev)
    e_off, e_cap = i64(np.arange(n, dtype=np.uint64) * np.uint64(estride)), i64(np.full(n, estride, dtype=np.uint64))
    e_size = torch.zeros(n, dtype=torch.int64, device=dev)
    e_st = torch.zeros(n, dtype=torch.int32, device=dev)
    wd = torch.full((n,), w, dtype=torch.int32, device=dev)
    ht = torch.full((n,), h, dtype=torch.int32, device=dev)

    def dec():
        ctx.bmp_decode_device(d_in, in_off, in_size, d_out, out_off, out_cap, out_size, st, stream=stream)

    def enc():
        ctx.bmp_encode_device(d_out, out_off, out_cap, wd, ht, d_enc, e_off, e_cap, e_size, e_st, stream=stream)

    res = {}
    for name, fn in (("decode", dec), ("encode", enc)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 10
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / steps
    assert int(st.abs().sum().item()) == 0 and int(e_st.abs().sum().item()) == 0, "bmp failures"
    exp = torch.from_numpy(np.frombuffer(rgba, np.uint8).copy()).to(dev)
    assert torch.equal(d_out.view(n, osz)[0], exp) and torch.equal(d_out.view(n, osz)[n - 1], exp), "bmp pixel mismatch"
    # the encoder's output is checked without the oracle (bench.py may only time it, in the CPU legs): 54 header
    # bytes as decode_bmp.c:313-358 writes them, then the pixels as BGRA, and 54 + size + 1 reported
    import struct
    size = 54 + osz + 1
    got = d_enc.view(n, estride)[n - 1][: size - 1].cpu().numpy().tobytes()
    hdr = b"BM" + struct.pack("<IHHI", (w * h * 4 + 54) & 0xffffffff, 0, 0, 54) + struct.pack(
        "<IiiHHIIIIII", 40, w, -h, 1, 32, 0, w * h * 4, 0, 0, 0, 0)
    bgra = np.frombuffer(rgba, np.uint8).reshape(-1, 4)[:, [2, 1, 0, 3]].tobytes()
    assert got == hdr + bgra and int(e_size[0].item()) == size, "bmp encode mismatch"
    alg = 2 * n * osz  # pixels read once + written once
    ach = alg / (res["decode"] / 1e3) / 1e9
    return {"metric": "bmp_decode_Mpixels_per_s", "value": n * w * h / (res["decode"] / 1e3) / 1e6, "unit": "Mpix/s",
            "ms_per_step": res["decode"], "encode_Mpixels_per_s": n * w * h / (res["encode"] / 1e3) / 1e6,
            "encode_ms_per_step": res["encode"],
            "config": {"workload": f"{n} x {w}x{h} 32-bit bottom-up BMP files (rows flipped), then re-encoded", "unique_images": 1},
            "roofline": {"bound": "hbm", "achieved": ach, "unit": "GB/s", "kernel": "bmp_swizzle_kernel",
                         "algorithmic_bytes_per_launch": alg, "encode_achieved": alg / (res["encode"] / 1e3) / 1e9}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=N_MEMBERS)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--png", type=int, default=1)
    ap.add_argument("--png-only", action="store_true")
    ap.add_argument("--png-images", type=int, default=PNG_N)
    ap.add_argument("--png-large", type=int, default=4, help="number of 8192x8192 images in the config-4-shape line (0 = skip)")
    ap.add_argument("--cfg5", type=int, default=2048, help="members of the config-5-shape line (0 = skip)")
    ap.add_argument("--bmp", type=int, default=64, help="number of 2048x2048 BMP files in the BMP line (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
