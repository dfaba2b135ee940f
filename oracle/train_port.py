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


def joints_shared_step(state_dict, imu, lengths, target, mask=None, prefix='joints.'):
    """-> (loss, {name: grad}) for the 20 tensors of one RNN head given as a state_dict with `prefix`."""
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
    loss = F.mse_loss(pred, target.view(target.shape[0], target.shape[1], -1)) + T_WEIGHT * temporal_loss(pred)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in sd.items()}, pred.detach()
