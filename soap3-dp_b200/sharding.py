"""Multi-GPU work partitioning: read (pair) batches are sharded across ranks,
the index is replicated in every GPU's HBM, and no data-path collective exists
(the reference's own recipe is "one process per GPU with -c <id>",
README.md:523-536).  Results come back to rank 0 in input order."""
from __future__ import annotations

from typing import List, Tuple


def shard_ranges(num_units: int, world_size: int, unit_multiple: int = 32) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) unit ranges per rank.  A unit is a read (single-end)
    or a pair (paired-end: the two mates are never split, rescue and pairing need
    both).  Boundaries fall on multiples of ``unit_multiple`` (the 32-read interleave
    of the query/answer buffers, QueryParser.cpp:1146-1152) so every shard can be cut
    out of the host buffers without repacking."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    blocks = (num_units + unit_multiple - 1) // unit_multiple
    out = []
    begin_blk = 0
    for r in range(world_size):
        nb = blocks // world_size + (1 if r < blocks % world_size else 0)
        b, e = begin_blk * unit_multiple, min((begin_blk + nb) * unit_multiple, num_units)
        out.append((min(b, num_units), e))
        begin_blk += nb
    return out


def gather_to_rank0(local, group=None):
    """Gather per-rank numpy arrays (first dimension = units of this shard) on rank 0
    and concatenate them in rank order == input order.  Uses torch.distributed
    (gloo on CPU tests, nccl or gloo on the GPU box); returns None on other ranks."""
    import numpy as np
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate(bucket, axis=0)
