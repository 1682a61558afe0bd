"""Host-side partitioning of a batch of independent streams / images across GPUs.

The path has no exchange step, so this is the whole multi-GPU story: the batch is cut into runs of consecutive items,
the runs are dealt to the devices longest-processing-time first by estimated decode time, every device decodes its
runs with its own context, results are gathered by item index. No collective touches the data path.

The logic lives in the library (dbg_multi_partition, csrc/dbg_api.cu), where dbg_decode_batch_packed_multi uses it with
one host thread per device; this module is the same function for callers that run one PROCESS per GPU (torchrun):
rank r keeps the items whose device index is r.
"""
import numpy as np

from . import api


def lpt_partition(in_size, world, out_cap=None, kind=0):
    """Returns `world` lists of item indices (index order inside a list). Items are taken to lie one after the other in
    both arenas, as the packed API wants them."""
    in_size = np.asarray(in_size, dtype=np.uint64)
    n = len(in_size)
    out_cap = np.asarray(out_cap if out_cap is not None else in_size, dtype=np.uint64)
    if n == 0:
        return [[] for _ in range(world)]
    in_off = np.concatenate([[0], np.cumsum(in_size[:-1] + np.uint64(16))]).astype(np.uint64)
    out_off = np.concatenate([[0], np.cumsum(out_cap[:-1])]).astype(np.uint64)
    dev, _, _ = api.partition(world, kind, in_off, in_size, out_off, out_cap)
    return [[int(i) for i in np.nonzero(dev == r)[0]] for r in range(world)]


def schedule_order(sizes):
    """Largest-first scheduling permutation for the per-GPU work queue."""
    return sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
