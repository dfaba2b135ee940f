"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Joints head's training step (first slice).

    joints_shared_step   mobileposer/models/joints.py:54-75  (MSE to the target joints + 1e-5 x temporal L1 of the second
                         differences, over the PADDED [B, T, 72] prediction exactly as the reference computes it)
    through RNN.forward  mobileposer/models/rnn.py:20-33     (dropout applied as an explicit mask so that a run is reproducible:
                         mask = None is eval mode / p = 0, otherwise the [B, T, H] tensor that multiplies relu(linear1(x)))

Like the reference it leans on torch (nn.LSTM / nn.Linear forward, autograd backward); pinned to the live reference's own
`Joints.shared_step` + `loss.backward()` by oracle/make_golden_train.py -> tests/golden/train_joints_step.npz.  Never imported by
the product package."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

T_WEIGHT = 1e-5      # joints.py:33


def temporal_loss(pred):
    """joints.py:72-75."""
    acc = pred[:, 2:, :] + pred[:, :-2, :] - 2 * pred[:, 1:-1, :]
    return torch.norm(acc, p=1, dim=2).sum(dim=1).mean()


def velocity_loss(pred, gt):
    """velocity.py:72-86: sum over n in {1, 3, 9} of the MSE of the T // n windows of n frames."""
    T = pred.shape[1]
    loss = 0.0
    for n in (1, 3, 9):
        for m in range(0, T // n):
            end = min(n * m + n, T)
            loss = loss + F.mse_loss(pred[:, m * n:end, :], gt[:, m * n:end, :])
    return loss


def jerk_loss(pred):
    """poser.py:100-103."""
    jerk = pred[:, 3:, :] - 3 * pred[:, 2:-1, :] + 3 * pred[:, 1:-2, :] - pred[:, :-3, :]
    return torch.norm(jerk, p=1, dim=2).sum(dim=1).mean()


def zero_pose_joints(pose_local):
    """Joint positions of the mean shape for local rotations [N, 24, 3, 3], root at the origin (articulate/model.py:208-232)."""
    from mobileposer_b200.config import SMPL_J_ZERO, SMPL_PARENT
    j = torch.tensor(SMPL_J_ZERO, dtype=pose_local.dtype)
    glb, pos = [pose_local[:, 0]], [torch.zeros(pose_local.shape[0], 3, dtype=pose_local.dtype)]
    for i in range(1, 24):
        p = SMPL_PARENT[i]
        glb.append(glb[p] @ pose_local[:, i])
        pos.append(pos[p] + glb[p] @ (j[i] - j[p]))
    return torch.stack(pos, dim=1)


def poser_loss(pred, pose_t96, joints_t):
    """poser.py:87-96 given the reduced target pose [B,T,96] and the target joints [B,T,72]."""
    from oracle.torch_port import reduced_global_to_full
    B, S = pred.shape[0], pred.shape[1]
    loss = F.mse_loss(pred, pose_t96) + T_WEIGHT * jerk_loss(pred)
    joints_p = zero_pose_joints(reduced_global_to_full(pred)).view(B, S, -1)
    return loss + F.mse_loss(joints_p, joints_t)


def head_shared_step(state_dict, x, lengths, target, kind, mask=None, prefix=''):
    """shared_step + backward of a head given its loss: kind = 'joints' (joints.py:54-75), 'footcontact' (footcontact.py:43-65, x =
    cat(noisy joints, imu)), 'velocity' (velocity.py:50-86).  -> (loss, {name: grad}, pred)."""
    loss_fn = {'joints': lambda p, t: F.mse_loss(p, t) + T_WEIGHT * temporal_loss(p),
               'footcontact': F.binary_cross_entropy_with_logits, 'velocity': velocity_loss,
               'poser': lambda p, t: poser_loss(p, t[..., :96], t[..., 96:])}[kind]      # target = cat(pose_t96, joints_t)
    return _shared_step(state_dict, x, lengths, target, mask, prefix, loss_fn)


def joints_shared_step(state_dict, imu, lengths, target, mask=None, prefix='joints.'):
    """-> (loss, {name: grad}) for the 20 tensors of one RNN head given as a state_dict with `prefix`."""
    return _shared_step(state_dict, imu, lengths, target, mask, prefix, lambda p, t: F.mse_loss(p, t) + T_WEIGHT * temporal_loss(p))


def _shared_step(state_dict, imu, lengths, target, mask, prefix, loss_fn):
    sd = {k[len(prefix):]: v.detach().clone().float().requires_grad_(True) for k, v in state_dict.items() if k.startswith(prefix)}
    hidden = sd['linear1.weight'].shape[0]
    bidir = 'rnn.weight_ih_l0_reverse' in sd
    lstm = torch.nn.LSTM(hidden, hidden, num_layers=2, bidirectional=bidir)
    a = torch.relu(F.linear(imu, sd['linear1.weight'], sd['linear1.bias']))
    if mask is not None:
        a = a * mask
    packed = pack_padded_sequence(a, lengths, batch_first=True, enforce_sorted=False)
    params = {k[4:]: v for k, v in sd.items() if k.startswith('rnn.')}
    out, _ = torch.func.functional_call(lstm, params, (packed,))
    out, _ = pad_packed_sequence(out, batch_first=True)
    pred = F.linear(out, sd['linear2.weight'], sd['linear2.bias'])
    loss = loss_fn(pred, target.view(target.shape[0], target.shape[1], -1))
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in sd.items()}, pred.detach()


def overfit_loop(state_dict, imu, lengths, target, mask, steps, prefix='joints.', kind='joints', lr=1e-3, gradient_clip_val=1.0):
    """The loop Lightning's Trainer(overfit_batches=1, gradient_clip_val=1) runs around training_step (overfit.py:41-56) with the
    optimizer of configure_optimizers (joints.py:113-114: torch.optim.AdamW(lr=1e-3)): zero_grad, shared_step, backward,
    clip_grad_norm_, step -- on ONE fixed batch.  -> (losses [steps], final {name: tensor})."""
    sd = {k[len(prefix):]: v.detach().clone().float() for k, v in state_dict.items() if k.startswith(prefix)}
    params = {k: torch.nn.Parameter(v) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=lr)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        loss, grads, _ = head_shared_step({k: v.detach() for k, v in params.items()}, imu, lengths, target, kind, mask)
        for k, p in params.items():
            p.grad = grads[k]
        if gradient_clip_val is not None:
            torch.nn.utils.clip_grad_norm_(list(params.values()), gradient_clip_val)
        opt.step()
        losses.append(loss)
    return torch.stack(losses), {k: v.detach() for k, v in params.items()}
