"""CPU suite: host-side sharding logic, single process and world_size-2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mobileposer_b200.sharding import gather_rows, shard_sequences


def test_shards_partition_and_balance():
    lengths = [3000] * 50
    for world in (1, 2, 4, 8):
        shards = shard_sequences(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(50))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= 3000
    ragged = [17, 3000, 250, 999, 4, 1200, 1200, 31, 640]
    shards = shard_sequences(ragged, 3)
    assert sorted(i for s in shards for i in s) == list(range(len(ragged)))
    loads = [sum(ragged[i] for i in s) for s in shards]
    assert max(loads) == 3000          # the longest sequence alone bounds the makespan
    assert shard_sequences(ragged, 3) == shards   # deterministic
    assert shard_sequences([], 2) == [[], []]
    with pytest.raises(ValueError):
        shard_sequences([1], 0)


def test_gather_rows_single_process():
    rows = torch.arange(12, dtype=torch.float32).view(4, 3)
    out = gather_rows(rows, [2, 0, 3, 1], 4)
    assert torch.equal(out[[2, 0, 3, 1]], rows)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lengths, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        shards = shard_sequences(lengths, world)
        mine = shards[rank]
        # per-sequence "metric rows": a deterministic function of the sequence id and length only
        rows = torch.stack([torch.tensor([float(i), float(lengths[i]), float(i) * 0.5]) for i in mine]) if mine \
            else torch.zeros(0, 3)
        out = gather_rows(rows, mine, len(lengths))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_gather_rows_gloo_is_shard_count_invariant(world):
    lengths = [300, 120, 3000, 45, 999, 300, 7]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lengths, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.tensor([[float(i), float(L), i * 0.5] for i, L in enumerate(lengths)])
    for _, out in results:
        assert torch.equal(out, expect)     # identical on every rank, identical to the 1-rank answer
