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
        out = gather_rows(rows, shards, len(lengths))
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


# ---- the whole evaluate_pose entry over two ranks (stand-in model on the CPU) ---------------------------------------
def _eval_items():
    g = torch.Generator().manual_seed(21)
    items = []
    for n in (70, 64, 90, 75, 66):
        pose_r6d = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24) + 0.2 * torch.randn(n, 144, generator=g)
        items.append((torch.randn(n, 60, generator=g), pose_r6d, torch.zeros(n, 24, 3), torch.cumsum(0.01 * torch.randn(n, 3, generator=g), 0)))
    return items


def _eval_worker(rank, world, port, batch_size, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mobileposer_b200.evaluate import evaluate_pose
        from test_evaluate import _StandInNet
        net = _StandInNet()
        table = evaluate_pose(net, _eval_items(), verbose=False, batch_size=batch_size)
        q.put((rank, table, net.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('batch_size', [1, 2])
def test_evaluate_pose_sharded_over_two_ranks_equals_one_rank(batch_size):
    """evaluate_pose under torch.distributed (gloo, world size 2): every rank evaluates its shard of whole sequences (looped
    or in batches), the [n, 8, 2] rows are all-gathered once, and every rank ends with the single-process table."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from mobileposer_b200.evaluate import evaluate_pose
    from test_evaluate import _StandInNet
    ref = evaluate_pose(_StandInNet(), _eval_items(), verbose=False)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, batch_size, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = []
    for rank, table, calls in results:
        assert torch.equal(torch.isnan(table), torch.isnan(ref))
        ok = ~torch.isnan(ref)
        assert torch.allclose(table[ok], ref[ok], rtol=1e-5, atol=1e-7)
        seen += [n for c in calls for n in c]
        assert all(len(c) <= batch_size for c in calls)
    assert sorted(seen) == sorted([70, 64, 90, 75, 66])       # every sequence evaluated exactly once across the ranks


# ---- a STATEFUL stand-in: the velocity-state carry must not make the table depend on the sharding ---------------
def _stateful_net():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_evaluate import _StandInNet

    class _Velocity:
        rnn_state = None

    class Stateful(_StandInNet):
        """forward_offline drifts by whatever `velocity.rnn_state` the previous B == 1 call left (like the reference's velocity
        head), and leaves its own: any result that depends on call order shows up in the Distance / Jitter rows."""

        def __init__(self):
            super().__init__()
            self.velocity = _Velocity()
            self.dynamics_optimizer = None

        def forward_offline(self, x, lengths):
            out = list(super().forward_offline(x, lengths))
            if x.shape[0] == 1:
                carry = self.velocity.rnn_state if self.velocity.rnn_state is not None else 0.0
                out[2] = out[2] + carry * torch.linspace(0, 1, out[2].shape[0]).view(-1, 1)
                self.velocity.rnn_state = carry + 0.05 * float(x.abs().mean())
            return tuple(out)
    return Stateful()


def _stateful_worker(rank, world, port, batch_size, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mobileposer_b200.evaluate import evaluate_pose
        q.put((rank, evaluate_pose(_stateful_net(), _eval_items(), verbose=False, batch_size=batch_size)))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_does_not_depend_on_the_carried_state():
    from mobileposer_b200.evaluate import evaluate_pose
    chained = evaluate_pose(_stateful_net(), _eval_items(), verbose=False)                 # one process, batch 1: the reference's chain
    independent = evaluate_pose(_stateful_net(), _eval_items(), verbose=False, batch_size=2)   # fresh state per sequence (5 = 2 + 2 + 1)
    ok = ~torch.isnan(chained)
    assert not torch.allclose(chained[ok], independent[ok], rtol=1e-4, atol=1e-6)          # the stand-in's leak is visible
    for world, batch_size in ((2, 1), (2, 2), (3, 1)):
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_stateful_worker, args=(r, world, port, batch_size, q)) for r in range(world)]
        for p in procs:
            p.start()
        results = [q.get(timeout=180) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for _, table in results:
            assert torch.allclose(table[ok], independent[ok], rtol=1e-5, atol=1e-7), (world, batch_size)


def _grad_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mobileposer_b200.training import average_gradients
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)       # this rank's gradient shard contribution
        scale = average_gradients(flat)
        q.put((rank, flat * scale))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_average_gradients_gloo(world):
    """The data-parallel exchange of a training step (training.average_gradients): one all-reduce of the flat gradient buffer; the
    returned factor turns the sum into the mean, identical on every rank."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    want = torch.arange(10, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    for r in range(world):
        assert torch.allclose(outs[r], want)


def test_average_gradients_without_a_process_group_is_the_identity():
    from mobileposer_b200.training import average_gradients
    flat = torch.ones(5)
    assert average_gradients(flat) == 1.0 and torch.equal(flat, torch.ones(5))
