"""Sequence sharding for multi-GPU evaluation.

Whole sequences are the unit (an LSTM pass is never split in time).  Ranks get a length-balanced partition
(greedy longest-first), run their shard with a full weight replica, and the only exchange of the path is one
all-gather of the per-sequence metric rows at the end.  Works with any torch.distributed backend (NCCL on the
GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_sequences(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Deterministic greedy bin packing by frame count; returns, per rank, the sorted sequence indices."""
    if world_size <= 0:
        raise ValueError('world_size must be positive')
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(s) for s in shards]


def gather_rows(local_rows: torch.Tensor, shards, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-sequence rows [n_local, ...] into dataset order [n_total, ...] on every rank.

    `shards` is the whole partition (`shard_sequences(lengths, world)`: every rank computes the same one from the lengths it
    already has), so neither ids nor counts travel: the exchange is ONE `all_gather_into_tensor` of a fixed
    [max shard size, ...] block per rank.  A flat list of ids is accepted for a single process.
    Rows of sequences nobody owned stay NaN."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if len(shards) == 0 or not isinstance(shards[0], (list, tuple)):
        shards = [list(shards)]
    if len(shards) != world:
        raise ValueError(f'gather_rows: partition has {len(shards)} shards for a world of {world}')
    if local_rows.shape[0] != len(shards[rank]):
        raise ValueError(f'gather_rows: rank {rank} holds {local_rows.shape[0]} rows for a shard of {len(shards[rank])}')
    feat = tuple(local_rows.shape[1:])
    dev = local_rows.device
    out = torch.full((n_total,) + feat, float('nan'), dtype=local_rows.dtype, device=dev)
    if world == 1:
        out[torch.as_tensor(shards[0], dtype=torch.int64, device=dev)] = local_rows
        return out
    n_max = max(1, max(len(s) for s in shards))
    block = torch.zeros((n_max,) + feat, dtype=local_rows.dtype, device=dev)
    block[:local_rows.shape[0]] = local_rows
    gathered = torch.empty((world * n_max,) + feat, dtype=local_rows.dtype, device=dev)
    dist.all_gather_into_tensor(gathered, block, group=group)
    gathered = gathered.view((world, n_max) + feat)
    for r, ids in enumerate(shards):
        if ids:
            out[torch.as_tensor(ids, dtype=torch.int64, device=dev)] = gathered[r, :len(ids)]
    return out
