"""Multi-GPU plumbing: one process per GPU, queries sharded by (map, goal) group, no data-path collective.

Queries never communicate.  The only exchanges are
  * broadcast_map: the N*N log-odds map (or the box list) from the rank that owns the update to the others
    (`ncclBroadcast` over NVLink on GPUs; gloo in the CPU tests), and
  * gather_records: fixed-size result records back to rank 0.
"""
import numpy as np


def shard_groups(n_groups_total, rank, world):
    """Round-robin partition of group ids over ranks: each rank builds / holds maps only for its groups."""
    return list(range(rank, n_groups_total, world))


def shard_queries(n_queries_total, rank, world):
    """Interleaved partition of query indices (every world-th query): with every map replicated on every rank this gives
    all ranks statistically equal work, whatever the difficulty of the individual groups (bench.py --workload c5)."""
    return np.arange(rank, n_queries_total, world)


def owner_of(group, world):
    return group % world


def broadcast_map(tensor, src, group=None):
    """In-place broadcast of one map tensor (float32, N*N) from rank `src`."""
    import torch.distributed as dist
    dist.broadcast(tensor, src=src, group=group)
    return tensor


def gather_records(local, world, group=None):
    """All ranks contribute a structured numpy array of fixed-size records (possibly different counts);
    returns the concatenation in rank order on every rank."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local)
    item = local.dtype.itemsize
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(local)], dtype=torch.int64), group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    buf = torch.zeros(max(m * item, 1), dtype=torch.uint8)
    raw = local.view(np.uint8).reshape(-1)
    buf[:len(raw)] = torch.from_numpy(raw.copy())
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    parts = [np.frombuffer(b.numpy().tobytes()[:c * item], dtype=local.dtype) for b, c in zip(bufs, counts)]
    return np.concatenate(parts) if parts else local
