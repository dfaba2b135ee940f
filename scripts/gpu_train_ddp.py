"""torchrun -N GPUs: data-parallel training steps of the Joints head (HeadTrainer + NCCL all-reduce of the flat gradient buffer) equal
the single-process steps on the whole batch.  Every rank holds the full replica, steps on its own B / N sequences, and ends with the
same parameters as one process stepping on all B.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/gpu_train_ddp.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobileposer_b200 as mp
from mobileposer_b200.synthetic import synthetic_imu_batch
from mobileposer_b200.training import HeadTrainer, dropout_mask

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
B, T, steps = 8 * world, 40, 4
gen = torch.Generator().manual_seed(11)
imu = synthetic_imu_batch(list(range(700, 700 + B)), T)
target = torch.randn(B, T, 72, generator=gen) * 0.3
mask = dropout_mask((B, T, 256), generator=gen)
lens = [T] * B


def run(group, lo, hi):
    torch.manual_seed(0)
    mod = mp.Joints().to(dev)
    tr = HeadTrainer(mod, process_group=group)
    losses = [tr.training_step(imu[lo:hi].to(dev), lens[lo:hi], target[lo:hi].to(dev), mask=mask[lo:hi].to(dev)).item() for _ in range(steps)]
    return losses, tr.flat_params.clone(), tr.grad_norm()


per = B // world
l_dp, p_dp, n_dp = run(None, rank * per, (rank + 1) * per)          # data parallel: this rank's shard, all-reduced gradients
l_one, p_one, n_one = run(False, 0, B)                              # one process, whole batch, no exchange
mean_loss = torch.tensor(l_dp, device=dev, dtype=torch.float64)
dist.all_reduce(mean_loss)
mean_loss /= world
err_p = (p_dp - p_one).abs().max().item()
err_l = (mean_loss.cpu() - torch.tensor(l_one, dtype=torch.float64)).abs().max().item()
gathered = [torch.zeros_like(p_dp) for _ in range(world)]
dist.all_gather(gathered, p_dp)
same = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    print(f'[ddp] {world} ranks x {per} sequences vs one process x {B}: max |param diff| {err_p:.2e}, max |mean loss diff| {err_l:.2e}, '
          f'grad norm {n_dp:.6f} vs {n_one:.6f}, replicas identical: {same}, flat buffer {p_dp.numel() * 4 / 1e6:.1f} MB')
    assert err_p < 5e-6 and err_l < 1e-6 and same
dist.barrier()
dist.destroy_process_group()
