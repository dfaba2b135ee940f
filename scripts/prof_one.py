"""Run a few eager (graph-free) forward_offline passes; the target of `ncu` captures (scripts/gpu_*.sh)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mobileposer_b200 as mp
from mobileposer_b200.synthetic import synthetic_imu_batch

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--frames', type=int, default=300)
ap.add_argument('--passes', type=int, default=2)
ap.add_argument('--tile', type=int, default=-1, help='>= 0: run through a HostOffline slot with this recurrence tile policy (bench.py uses 64)')
ap.add_argument('--physics', action='store_true', help='PHYSICS hook on (K8 at the end of every pass)')
a = ap.parse_args()

torch.manual_seed(0)
net = mp.MobilePoserNet().eval().to('cuda:0')
net.set_graph(False)
net.reuse_outputs = True
net.enable_physics(a.physics)
x = synthetic_imu_batch(list(range(a.batch)), a.frames).to('cuda:0')
slot = mp.HostOffline(net, a.batch, a.frames, rec_tile=a.tile) if a.tile >= 0 else None
if slot is not None:
    _ = slot.handle
    from mobileposer_b200 import _cabi
    _cabi.check(_cabi.lib().mp_net_set_graph(slot.handle, 0))
for _ in range(a.passes):
    if slot is not None:
        slot.submit_device(x)
        slot.wait()
        continue
    net.velocity.rnn_state = None
    net.forward_offline(x, [a.frames] * a.batch)
torch.cuda.synchronize()
print('done', a.batch, a.frames)
