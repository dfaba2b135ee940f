"""Time mp_physics_optimize on B x T synthetic frames (CUDA events, current stream)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from mobileposer_b200.dynamics import PhysicsOptimizer
from physics_inputs import synthetic_motion

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--frames', type=int, default=300)
ap.add_argument('--iters', type=int, default=10)
a = ap.parse_args()
R, vel, contact = synthetic_motion(a.batch, a.frames, seed=1)
tR, tv, tc = (torch.from_numpy(v).cuda() for v in (R, vel, contact))
opt = PhysicsOptimizer()
out = torch.empty(a.batch, a.frames, 24, 9, device='cuda')
for _ in range(2):
    opt.reset_states(); opt.optimize_sequences(tR, tv, tc, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    opt.optimize_sequences(tR, tv, tc, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
fr = a.batch * a.frames
print(f'k8 physics B={a.batch} T={a.frames}: {ms:.3f} ms/launch, {ms * 1e3 / a.frames:.2f} us/frame-step, {fr / ms * 1e3:.3e} frames/s, '
      f'{fr * 2036 / ms / 1e6:.1f} GB/s algorithmic')
