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
    class FixedMask(torch.nn.Module):        # stands in for nn.Dropout(p=0.4): the same keep / (1 - p) scaling, a fixed pattern
        def forward(self, x, m=keep):
            return x * m

    for tag, mask in (('eval', None), ('mask', keep)):
        mod.zero_grad()
        if mask is None:
            mod.eval()
        else:
            mod.train()
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
    print('oracle/train_port.py agrees with the live reference (joints)')

    # ---- the optimisation loop: Lightning is not installed here, so its Trainer(overfit_batches=1, gradient_clip_val=c) loop is spelled
    # out around the LIVE module's own shared_step and configure_optimizers (AdamW, lr 1e-3): zero_grad, shared_step, backward,
    # clip_grad_norm_, step -- 6 steps on the fixed batch above, dropout as the fixed mask, c = 1 (overfit.py:46) and c = 0.005 (so that
    # the clip is active) ------------------------------------------------------------------------------------------------------------
    from oracle.train_port import overfit_loop
    out3 = {'imu': imu, 'lengths': np.asarray(lens), 'target': target, 'mask': keep}
    for tag, clip in (('clip1', 1.0), ('clip005', 0.005)):
        os.chdir(os.path.join(REF, 'mobileposer'))
        try:
            torch.manual_seed(0)
            mod = Joints()
        finally:
            os.chdir(cwd)
        mod.train()
        mod.joints.dropout = FixedMask()
        sd0 = {k: v.detach().clone() for k, v in mod.state_dict().items() if k.startswith('joints.')}
        opt = mod.configure_optimizers()
        losses, norms = [], []
        for _ in range(6):
            opt.zero_grad()
            loss = mod.shared_step(((imu, lens), ({'joints': target}, None)))
            loss.backward()
            norms.append(torch.nn.utils.clip_grad_norm_(mod.parameters(), clip))
            opt.step()
            losses.append(loss.detach())
        out3[f'{tag}_losses'] = torch.stack(losses)
        out3[f'{tag}_grad_norms'] = torch.stack(norms)
        for name, prm in mod.named_parameters():
            gname = name[len('joints.'):]
            val = prm.detach()
            out3[f'{tag}_pnorm.{gname}'] = val.norm()
            out3[f'{tag}_param.{gname}'] = val if val.numel() <= 40000 else val[::7, ::5].contiguous()
        p_losses, p_final = overfit_loop(sd0, imu, lens, target, keep, 6, gradient_clip_val=clip)
        assert (p_losses - out3[f'{tag}_losses']).abs().max() <= 1e-6 * out3[f'{tag}_losses'].abs().max(), (tag, p_losses, out3[f'{tag}_losses'])
        for k, v in p_final.items():
            assert abs(v.norm().item() - out3[f'{tag}_pnorm.{k}'].item()) <= 1e-5 * out3[f'{tag}_pnorm.{k}'].item(), (tag, k)
        print(tag, 'losses', [round(float(v), 6) for v in losses], 'grad norms', [round(float(v), 4) for v in norms])
    np.savez_compressed(os.path.join(OUT, 'train_overfit_joints.npz'), **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out3.items()})
    print('wrote train_overfit_joints.npz; oracle/train_port.py overfit_loop agrees with the live reference')

    # ---- FootContact (footcontact.py:43-65, BCE with logits) and Velocity (velocity.py:50-86, windowed MSE), eval mode; the noise the
    # reference draws inside shared_step (torch.randn right after torch.manual_seed) is reproduced and the noisy input saved --------------
    from oracle.train_port import head_shared_step
    os.chdir(os.path.join(REF, 'mobileposer'))
    try:
        from mobileposer.models import FootContact, Velocity
        torch.manual_seed(0)
        foot = FootContact()
        torch.manual_seed(0)
        velm = Velocity()
    finally:
        os.chdir(cwd)
    joints_gt = torch.randn(B, T, 24, 3, generator=g) * 0.3
    contacts = (torch.rand(B, T, 2, generator=g) > 0.5).float()
    vels = torch.randn(B, T, 24, 3, generator=g) * 0.5
    out2 = {'imu': imu, 'lengths': np.asarray(lens), 'foot_contacts': contacts, 'vels': vels}
    for name, mod, std, outputs, kind, prefix in (('foot', foot, 0.04, {'joints': joints_gt.clone(), 'foot_contacts': contacts}, 'footcontact', 'footcontact.'),
                                                  ('vel', velm, 0.025, {'joints': joints_gt.clone(), 'vels': vels}, 'velocity', 'vel.')):
        mod.eval()
        mod.zero_grad()
        torch.manual_seed(77)
        noise = torch.randn(B, T, 72) * std
        torch.manual_seed(77)                       # shared_step draws the same tensor
        loss = mod.shared_step(((imu, lens), (outputs, None)))
        loss.backward()
        x_cat = torch.cat((joints_gt.view(B, T, 72) + noise, imu), dim=-1)
        out2[f'{name}_input'] = x_cat
        out2[f'{name}_loss'] = loss.detach()
        sdm = {k: v.detach().clone() for k, v in mod.state_dict().items() if k.startswith(prefix)}
        target = contacts if name == 'foot' else vels.view(B, T, 72)
        p_loss, p_grads, _ = head_shared_step(sdm, x_cat, lens, target, kind, prefix=prefix)
        assert abs(p_loss.item() - loss.item()) < 1e-6 * max(1.0, abs(loss.item())), (name, p_loss, loss)
        for pname, prm in mod.named_parameters():
            if not pname.startswith(prefix):
                continue
            gname = pname[len(prefix):]
            gr = prm.grad.detach()
            assert (p_grads[gname] - gr).abs().max() <= 1e-6 * max(1.0, gr.abs().max().item()), (name, gname)
            out2[f'{name}_norm.{gname}'] = gr.norm()
            out2[f'{name}_grad.{gname}'] = gr if gr.numel() <= 40000 else gr[::7, ::5].contiguous()
    # ---- Poser (poser.py:65-98): MSE + jerk L1 + joint-position loss through _reduced_global_to_full and the body model's FK --------
    os.chdir(os.path.join(REF, 'mobileposer'))
    try:
        from mobileposer.models import Poser
        torch.manual_seed(0)
        poser = Poser()
    finally:
        os.chdir(cwd)
    eye6 = torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(24)
    poses = eye6 + 0.3 * torch.randn(B, T, 144, generator=g)                 # target pose: 24 joints x r6d
    poser.eval()
    poser.zero_grad()
    torch.manual_seed(78)
    noise = torch.randn(B, T, 72) * 0.04
    torch.manual_seed(78)
    loss = poser.shared_step(((imu, lens), ({'poses': poses, 'joints': joints_gt.clone()}, None)))
    loss.backward()
    x_cat = torch.cat((joints_gt.view(B, T, 72) + noise, imu), dim=-1)
    import mobileposer.config as RC
    pose_t96 = poses.view(B, T, 24, 6)[:, :, RC.joint_set.reduced].reshape(B, T, 96)
    out2.update(pose_input=x_cat, poses=poses, joints_gt=joints_gt.view(B, T, 72), pose_loss=loss.detach())
    sdm = {k: v.detach().clone() for k, v in poser.state_dict().items() if k.startswith('pose.')}
    p_loss, p_grads, _ = head_shared_step(sdm, x_cat, lens, torch.cat((pose_t96, joints_gt.view(B, T, 72)), dim=-1), 'poser', prefix='pose.')
    assert abs(p_loss.item() - loss.item()) < 1e-6 * max(1.0, abs(loss.item())), (p_loss, loss)
    for pname, prm in poser.named_parameters():
        if not pname.startswith('pose.'):
            continue
        gname = pname[len('pose.'):]
        gr = prm.grad.detach()
        assert (p_grads[gname] - gr).abs().max() <= 2e-6 * max(1.0, gr.abs().max().item()), ('pose', gname, (p_grads[gname] - gr).abs().max())
        out2[f'pose_norm.{gname}'] = gr.norm()
        out2[f'pose_grad.{gname}'] = gr if gr.numel() <= 40000 else gr[::7, ::5].contiguous()
    print('oracle/train_port.py agrees with the live reference (poser)')
    np.savez_compressed(os.path.join(OUT, 'train_heads_step.npz'), **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out2.items()})
    print('wrote train_heads_step.npz; oracle/train_port.py agrees with the live reference (footcontact, velocity)')


if __name__ == '__main__':
    main()
