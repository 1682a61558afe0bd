"""Multi-GPU host logic on CPU: world_size-2 gloo ranks partition a batch with
no data-path collective and agree on the max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from debigulator_b200.shard import lpt_partition, schedule_order


def test_lpt_partition_properties():
    sizes = [(i * 7919) % 1000 + 1 for i in range(257)]
    for world in (1, 2, 4, 8):
        shards = lpt_partition(sizes, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in s) for s in shards]
        # runs of ~1/8 of a device's share, dealt longest first: the loads differ by less than one run
        assert max(loads) - min(loads) <= sum(sizes) / (8 * world) + 2 * max(sizes) + 8 * 4096, (world, loads)
    assert lpt_partition([], 2) == [[], []]
    order = schedule_order(sizes)
    assert [sizes[i] for i in order] == sorted(sizes, reverse=True)


def test_partition_few_huge_items():
    """BASELINE config 4 shape: 256 equal images over 8 devices -> 32 each; and one dominant item gets a device alone."""
    shards = lpt_partition([150_000_000] * 256, 8, out_cap=[268_435_456] * 256, kind=2)
    assert sorted(len(s) for s in shards) == [32] * 8
    shards = lpt_partition([10_000_000] + [100_000] * 64, 2)
    big = [s for s in shards if 0 in s][0]
    assert len(big) < 20


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [(i * 31) % 97 + 1 for i in range(100)]
    mine = lpt_partition(sizes, world)[rank]
    # every rank "decodes" its shard: here, a checksum of item ids stands in for the per-item results
    local = torch.zeros(len(sizes), dtype=torch.int64)
    for i in mine:
        local[i] = sizes[i] * 3 + 1
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)          # result gathering only; the data path itself has no collective
    total = torch.stack(gathered).sum(0)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing
    q.put((rank, total.tolist(), float(t.item()), len(mine)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    sizes = [(i * 31) % 97 + 1 for i in range(100)]
    for rank, total, tmax, n_mine in res:
        assert total == [x * 3 + 1 for x in sizes]
        assert tmax == 2.0
        assert 40 <= n_mine <= 60
