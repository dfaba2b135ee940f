"""Training-step timing at the reference's training shape (train.py / data.py:148-165: batch 256 windows x 125 frames): every head's
shared_step + backward + clip + AdamW on the device (HeadTrainer), CUDA events; the Joints head beside the CPU oracle (torch autograd,
all host threads) on the same batch."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mobileposer_b200 as mp
from mobileposer_b200 import _cabi
from mobileposer_b200.synthetic import synthetic_imu_batch
from mobileposer_b200.training import HeadTrainer, dropout_mask

dev = 'cuda:0'
B, T = int(os.environ.get('B', 256)), int(os.environ.get('T', 125))
gen = torch.Generator().manual_seed(1)
imu = synthetic_imu_batch(list(range(B)), T)
lens = [T] * B
joints_t = torch.randn(B, T, 72, generator=gen) * 0.3
eye6 = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24)
poses = eye6 + 0.3 * torch.randn(B, T, 144, generator=gen)
cat_in = torch.cat((joints_t + 0.04 * torch.randn(B, T, 72, generator=gen), imu), -1)
contacts = (torch.rand(B, T, 2, generator=gen) > 0.5).float()
vels = torch.randn(B, T, 72, generator=gen) * 0.5
cases = [('joints', mp.Joints, (imu, lens, joints_t), 256), ('poser', mp.Poser, (cat_in, lens, poses, joints_t), 256),
         ('footcontact', mp.FootContact, (cat_in, lens, contacts), 64), ('velocity', mp.Velocity, (cat_in, lens, vels), 256)]
for name, cls, args, H in cases:
    torch.manual_seed(0)
    tr = HeadTrainer(cls().to(dev))
    dargs = tuple(a.to(dev) if torch.is_tensor(a) else a for a in args)
    mask = dropout_mask((B, T, H), generator=gen).to(dev)
    for _ in range(3):
        loss = tr.training_step(*dargs, mask=mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        loss = tr.training_step(*dargs, mask=mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f'[train time] {name}: {ms:.2f} ms per step (B={B}, T={T}: {B * T / ms / 1e3:.2f} M frames/s), loss {loss.item():.5f}, '
          f'{tr.flat_params.numel() * 4 / 1e6:.1f} MB of parameters')

# CPU oracle, Joints head, same batch
from oracle.train_port import overfit_loop
torch.manual_seed(0)
sd = {'joints.' + k: v for k, v in mp.Joints().joints.state_dict().items()}
mask = dropout_mask((B, T, 256), generator=gen)
t0 = time.perf_counter()
steps = 2
overfit_loop(sd, imu, lens, joints_t, mask, steps)
dt = (time.perf_counter() - t0) / steps
print(f'[train time] CPU oracle (torch autograd + AdamW, {torch.get_num_threads()} threads), joints: {dt * 1e3:.0f} ms per step '
      f'({B * T / dt / 1e3:.1f} k frames/s)')
