"""cfg4 (BASELINE.json config 4): the evaluate.py-shaped run -- 10 subjects x 5 sequences x 3000 frames through
`evaluate_pose` (reset + forward_offline per sequence, evaluate.py:56-58, then the evaluator rows on the device: per-frame
errors, mesh row over a 6890-vertex template, translation windows).  Wall clock of the whole entry on one GPU, and of the
model calls alone.  The mesh template is synthetic (the SMPL file cannot travel); the ground truth is synthetic_dip's."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobileposer_b200.config import SMPL_J_ZERO
from mobileposer_b200.evaluate import evaluate_pose, synthetic_dip
from mobileposer_b200.net import MobilePoserNet

dev = 'cuda:0'
torch.manual_seed(0)
net = MobilePoserNet().to(dev).eval()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
t0 = time.perf_counter()
items = synthetic_dip(n_subjects=1, n_seq=5, frames=frames) * 10      # 5 distinct sequences, each evaluated 10 times: 50 items
print(f'synthetic set: {len(items)} sequences x {frames} frames generated in {time.perf_counter() - t0:.1f} s (host, untimed)')
g = torch.Generator().manual_seed(1)
V = 6890
jz = torch.tensor(SMPL_J_ZERO)
near = torch.randint(0, 24, (V,), generator=g)
rest = (jz[near] + torch.randn(V, 3, generator=g) * 0.05).to(dev)
w = torch.zeros(V, 24)
w[torch.arange(V), near] = 1.0
w.scatter_add_(1, torch.randint(0, 24, (V, 3), generator=g), torch.rand(V, 3, generator=g) * 0.5)
mesh = (rest, (w / w.sum(1, keepdim=True)).to(dev))

evaluate_pose(net, items[:2], evaluate_tran=True, mesh=mesh, verbose=False)      # warm-up: module load, graph capture
torch.cuda.synchronize()
for label, kw in (('evaluate_pose: model + all evaluator rows (mesh, translation windows)', dict(evaluate_tran=True, mesh=mesh)),
                  ('evaluate_pose: model + joint / angle rows (the default call)', dict())):
    t0 = time.perf_counter()
    out = evaluate_pose(net, items, verbose=False, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'{label}: {dt * 1e3:.0f} ms for {len(items) * frames} frames = {len(items) * frames / dt / 1e3:.0f} k frames/s, {dt / len(items) * 1e3:.1f} ms per sequence')
xs = [it[0].to(dev) for it in items]
torch.cuda.synchronize()
t0 = time.perf_counter()
for x in xs:
    net.reset()
    net.forward_offline(x.unsqueeze(0), [x.shape[0]])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'model calls alone (inputs resident): {dt * 1e3:.0f} ms = {len(items) * frames / dt / 1e3:.0f} k frames/s, {dt / len(items) * 1e3:.1f} ms per sequence')
