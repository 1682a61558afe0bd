"""Host-side partitioning of a batch of independent streams / images across
GPUs (one process per GPU). The path has no exchange step, so this is the whole
multi-GPU story: greedy longest-processing-time assignment by compressed size,
each rank decodes its own shard with its own context, results are gathered by
item index. No collective touches the data path."""
import heapq


def lpt_partition(sizes, world):
    """Returns `world` lists of item indices; heaviest items first, each to the
    currently lightest rank. Deterministic for equal inputs."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + int(sizes[i]), r))
    return shards


def schedule_order(sizes):
    """Largest-first scheduling permutation for the per-GPU work queue."""
    return sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
