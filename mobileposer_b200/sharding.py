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


def gather_rows(local_rows: torch.Tensor, local_ids: Sequence[int], n_total: int, group=None) -> torch.Tensor:
    """All-gather per-sequence rows [n_local, ...] into dataset order [n_total, ...] on every rank.

    One `all_gather_into_tensor` of a padded block plus the ids; rows of sequences nobody owned stay NaN."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    feat = tuple(local_rows.shape[1:])
    out = torch.full((n_total,) + feat, float('nan'), dtype=local_rows.dtype, device=local_rows.device)
    ids = torch.as_tensor(list(local_ids), dtype=torch.int64, device=local_rows.device)
    if world == 1:
        out[ids] = local_rows
        return out
    n_max = (n_total + world - 1) // world
    n_max = max(n_max, 1)
    counts = torch.tensor([len(local_ids)], dtype=torch.int64, device=local_rows.device)
    all_counts = torch.empty(world, dtype=torch.int64, device=local_rows.device)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    n_max = int(all_counts.max().item())
    pad_rows = torch.zeros((n_max,) + feat, dtype=local_rows.dtype, device=local_rows.device)
    pad_rows[:len(local_ids)] = local_rows
    pad_ids = torch.full((n_max,), -1, dtype=torch.int64, device=local_rows.device)
    pad_ids[:len(local_ids)] = ids
    g_rows = torch.empty((world * n_max,) + feat, dtype=local_rows.dtype, device=local_rows.device)
    g_ids = torch.empty(world * n_max, dtype=torch.int64, device=local_rows.device)
    dist.all_gather_into_tensor(g_rows, pad_rows, group=group)
    dist.all_gather_into_tensor(g_ids, pad_ids, group=group)
    keep = g_ids >= 0
    out[g_ids[keep]] = g_rows[keep]
    return out
