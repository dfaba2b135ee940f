"""Isolated timing of the evaluator kernels at DIP size: one 3000-frame sequence, a
6890-vertex template (synthetic: the SMPL file cannot travel), and 50 sequences of translation windows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobileposer_b200.config import SMPL_J_ZERO
from mobileposer_b200.evaluate import frame_errors_cuda, tran_window_errors, vertex_error_row

dev = 'cuda:0'
g = torch.Generator().manual_seed(0)
n, V, S = 3000, 6890, 50


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


q = torch.linalg.qr(torch.randn(2 * n * 24, 3, 3, generator=g))[0]
q = q * torch.linalg.det(q).sign().view(-1, 1, 1)
pose_p, pose_t = q[:n * 24].view(n, 24, 3, 3).to(dev), q[n * 24:].view(n, 24, 3, 3).to(dev)
jz = torch.tensor(SMPL_J_ZERO)
near = torch.randint(0, 24, (V,), generator=g)
rest = (jz[near] + torch.randn(V, 3, generator=g) * 0.05).to(dev)
w = torch.zeros(V, 24)
w[torch.arange(V), near] = 1.0
w.scatter_add_(1, torch.randint(0, 24, (V, 3), generator=g), torch.rand(V, 3, generator=g) * 0.5)
w = (w / w.sum(1, keepdim=True)).to(dev)
tran = torch.cumsum(torch.randn(S, n, 3, generator=g).abs() * 0.004, 1).to(dev)
tran_p = tran + torch.cumsum(torch.randn(S, n, 3, generator=g) * 0.001, 1).to(dev)

ms = timed(lambda: vertex_error_row(pose_p, pose_t, (rest, w)))
print(f'mesh row   n={n} V={V}: {ms:.3f} ms  ({n * V * 288 * 2 / ms / 1e9:.1f} TFLOP/s of skinning FMAs; the materialising formulation '
      f'would move {2 * n * V * 12 / 1e6:.0f} MB of vertices = {2 * n * V * 12 / 6.5e12 * 1e3:.3f} ms at the HBM roofline)')
ms = timed(lambda: frame_errors_cuda(pose_p, pose_t, tran[0], tran_p[0]))
print(f'frame rows n={n}: {ms:.3f} ms  ({n * (2 * 864 + 24 + 2 * 288 + 3 * 96) / ms / 1e6:.1f} GB/s algorithmic)')
ms = timed(lambda: tran_window_errors(tran_p, tran))
print(f'tran windows S={S} T={n}: {ms:.3f} ms  ({ms / S * 1e3:.1f} us per sequence; the sequential fp32 distance is the floor)')
