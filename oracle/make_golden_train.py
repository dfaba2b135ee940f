"""TEST INFRASTRUCTURE ONLY -- tests/golden/train_joints_step.npz from the LIVE reference: `Joints.shared_step` (joints.py:54-75) +
`loss.backward()` on a small ragged batch, once in eval mode (dropout off) and once with the module's dropout replaced by a fixed
seeded mask (so the run is reproducible).  Build container only (needs /root/reference).

    python oracle/make_golden_train.py
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shims'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import numpy as np
import torch

from mobileposer_b200.synthetic import synthetic_imu_batch

OUT = os.path.join(ROOT, 'tests', 'golden')


def main():
    torch.set_num_threads(1)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, 'mobileposer'))
    try:
        from mobileposer.models import Joints
        torch.manual_seed(0)
        mod = Joints()
    finally:
        os.chdir(cwd)
    lens = [20, 13, 17]
    B, T = 3, 20
    imu = synthetic_imu_batch([301, 302, 303], T)
    for b, L in enumerate(lens):
        imu[b, L:] = 0
    g = torch.Generator().manual_seed(91)
    target = torch.randn(B, T, 24, 3, generator=g) * 0.3
    keep = (torch.rand(B, T, 256, generator=g) >= 0.4).float() / 0.6         # nn.Dropout(p=0.4) in training mode: keep / (1 - p)
    out = {'imu': imu, 'lengths': np.asarray(lens), 'target': target, 'mask': keep}
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items() if k.startswith('joints.')}
    for tag, mask in (('eval', None), ('mask', keep)):
        mod.zero_grad()
        if mask is None:
            mod.eval()
        else:
            mod.train()
            class FixedMask(torch.nn.Module):        # stands in for nn.Dropout(p=0.4): the same keep / (1 - p) scaling, a fixed pattern
                def forward(self, x, m=mask):
                    return x * m
            mod.joints.dropout = FixedMask()
        batch = ((imu, lens), ({'joints': target}, None))
        loss = mod.shared_step(batch)
        loss.backward()
        out[f'{tag}_loss'] = loss.detach()
        for name, p in mod.named_parameters():
            if not name.startswith('joints.'):
                continue
            gname = name[len('joints.'):]
            gr = p.grad.detach()
            out[f'{tag}_norm.{gname}'] = gr.norm()
            # small tensors whole, the [1024, 256 | 512] LSTM matrices as a strided sample
            out[f'{tag}_grad.{gname}'] = gr if gr.numel() <= 40000 else gr[::7, ::5].contiguous()
    np.savez_compressed(os.path.join(OUT, 'train_joints_step.npz'), **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print('wrote train_joints_step.npz', {k: tuple(np.asarray(v).shape) for k, v in out.items() if 'grad' not in k and 'norm' not in k})
    # the port against what was just written
    from oracle.train_port import joints_shared_step
    for tag, mask in (('eval', None), ('mask', keep)):
        loss, grads, _ = joints_shared_step(sd, imu, lens, target, mask)
        assert abs(loss.item() - out[f'{tag}_loss'].item()) < 1e-6, (tag, loss, out[f'{tag}_loss'])
        for k, gr in grads.items():
            ref = out[f'{tag}_grad.{k}']
            got = gr if gr.numel() <= 40000 else gr[::7, ::5]
            assert (got - ref).abs().max() <= 1e-6 * max(1.0, ref.abs().max().item()), (tag, k)
    print('oracle/train_port.py agrees with the live reference')


if __name__ == '__main__':
    main()
