"""Training steps of the Joints, FootContact and Velocity heads -- the first slice of the reference's training path on the device
(SURVEY.md 8f row N4).

    joints_shared_step(module, imu, lengths, target, mask=None) -> (loss, grads, pred)
    footcontact_shared_step / velocity_shared_step: the same for footcontact.py:43-65 (BCE with logits) and velocity.py:50-86
    (windowed MSE); poser_shared_step: poser.py:65-98 (MSE + jerk L1 + the joint-position loss through _reduced_global_to_full and
    the zero-pose forward kinematics, with the Gram-Schmidt and tree adjoints in the loss kernel).

is `Joints.shared_step` (mobileposer/models/joints.py:54-75: MSE to the target joints + 1e-5 x temporal L1 of the second
differences) followed by `loss.backward()`: the forward of `RNN.forward` (models/rnn.py:20-33) with saved activations, the loss and
its gradient, and the backward pass through linear2, both LSTM layers and directions, the dropout mask, ReLU and linear1 -- all in
CUDA kernels behind the C ABI (mp_rnn_train_forward / mp_joints_loss / mp_rnn_train_backward), gradients in torch's own layouts under
the parameter names of the head.  `mask` stands in for nn.Dropout (the reference's p = 0.4 keep / (1 - p) pattern, drawn by the
caller so a step is reproducible); None = eval mode.  No optimizer, no other head yet: DESIGN.md section 7."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from .modules import RNN, _f32c, _require_cuda, current_stream_ptr

T_WEIGHT = 1e-5      # joints.py:33


def dropout_mask(shape, p=0.4, generator=None, device=None):
    """nn.Dropout(p) in training mode as an explicit tensor: keep / (1 - p) with probability 1 - p, else 0."""
    keep = torch.rand(shape, generator=generator, device=generator.device if generator is not None else device) >= p
    return (keep.float() / (1.0 - p)).to(device) if device is not None else keep.float() / (1.0 - p)


def _weights_struct(rnn: RNN, tensors):
    w = _cabi.RnnWeights()
    w.n_input, w.n_output, w.n_hidden = rnn.n_input, rnn.n_output, rnn.n_hidden
    w.n_layers, w.bidirectional = rnn.n_rnn_layer, int(rnn.bidirectional)
    w.linear1_w, w.linear1_b = tensors['linear1.weight'].data_ptr(), tensors['linear1.bias'].data_ptr()
    w.linear2_w, w.linear2_b = tensors['linear2.weight'].data_ptr(), tensors['linear2.bias'].data_ptr()
    for layer in range(2):
        for d in range(2 if rnn.bidirectional else 1):
            sfx = f'_l{layer}' + ('_reverse' if d else '')
            w.w_ih[layer][d] = tensors['rnn.weight_ih' + sfx].data_ptr()
            w.w_hh[layer][d] = tensors['rnn.weight_hh' + sfx].data_ptr()
            w.b_ih[layer][d] = tensors['rnn.bias_ih' + sfx].data_ptr()
            w.b_hh[layer][d] = tensors['rnn.bias_hh' + sfx].data_ptr()
    return w


@torch.no_grad()
def rnn_forward_backward(rnn: RNN, x, lengths, dloss_dy_fn, mask=None):
    """Forward of one head with saved activations, `dy = dloss_dy_fn(y)` on the device, backward.  -> (y, grads {param name: tensor})."""
    _require_cuda(x, 'training input')
    if rnn.n_rnn_layer != 2:
        raise ValueError('only the reference\'s 2-layer LSTM is built')
    lib = _cabi.lib()
    dev = x.device
    xa = _f32c(x)
    B, T = xa.shape[0], xa.shape[1]
    lens = [int(v) for v in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
    if len(lens) != B or min(lens) <= 0 or max(lens) != T:
        raise RuntimeError('lengths must have one positive entry per sequence and max(lengths) must equal the padded length')
    params = {k: _f32c(v) for k, v in rnn.named_parameters()}
    _require_cuda(params['linear1.weight'], 'RNN parameters')
    w = _weights_struct(rnn, params)
    m = _f32c(mask) if mask is not None else None
    if m is not None and tuple(m.shape) != (B, T, rnn.n_hidden):
        raise ValueError(f'dropout mask must be {(B, T, rnn.n_hidden)}, got {tuple(m.shape)}')
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=dev)
    y = torch.empty(B, T, rnn.n_output, device=dev, dtype=torch.float32)
    grads = {k: torch.empty_like(v) for k, v in params.items()}
    g = _cabi.RnnGrads()
    gw = _weights_struct(rnn, grads)
    for f, _ in _cabi.RnnGrads._fields_:
        setattr(g, f, getattr(gw, f))
    with torch.cuda.device(dev):
        stream = current_stream_ptr(dev)
        ws_bytes = lib.mp_rnn_train_workspace_bytes(C.byref(w), B, T)
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        mp = m.data_ptr() if m is not None else None
        _cabi.check(lib.mp_rnn_train_forward(C.byref(w), xa.data_ptr(), B, T, lens_dev.data_ptr(), mp, y.data_ptr(), ws.data_ptr(),
                                             ws_bytes, stream), 'mp_rnn_train_forward')
        dy = _f32c(dloss_dy_fn(y))
        _cabi.check(lib.mp_rnn_train_backward(C.byref(w), xa.data_ptr(), B, T, lens_dev.data_ptr(), mp, dy.data_ptr(), C.byref(g),
                                              ws.data_ptr(), ws_bytes, stream), 'mp_rnn_train_backward')
    return y, grads


@torch.no_grad()
def joints_shared_step(module, imu, lengths, target, mask=None):
    """Joints.shared_step + backward (joints.py:54-75).  module: mobileposer_b200.Joints on a CUDA device; imu [B,T,60];
    target [B,T,24,3] or [B,T,72].  -> (loss tensor [] float64 on the device, {'joints.<param>': grad}, pred [B,T,72])."""
    lib = _cabi.lib()
    box = {}

    def dloss(pred):
        B, T, D = pred.shape
        tgt = _f32c(target.to(pred.device)).view(B, T, D)
        loss = torch.zeros((), device=pred.device, dtype=torch.float64)
        dpred = torch.empty_like(pred)
        _cabi.check(lib.mp_joints_loss(pred.data_ptr(), tgt.data_ptr(), B, T, D, T_WEIGHT, loss.data_ptr(), dpred.data_ptr(),
                                       current_stream_ptr(pred.device)), 'mp_joints_loss')
        box['loss'] = loss
        return dpred

    pred, grads = rnn_forward_backward(module.joints, imu, lengths, dloss, mask)
    return box['loss'], {'joints.' + k: v for k, v in grads.items()}, pred


def _step_with_loss(rnn, prefix, x, lengths, target, mask, loss_call):
    lib = _cabi.lib()
    box = {}

    def dloss(pred):
        B, T, D = pred.shape
        tgt = _f32c(target.to(pred.device)).view(B, T, D)
        loss = torch.zeros((), device=pred.device, dtype=torch.float64)
        dpred = torch.empty_like(pred)
        loss_call(lib, pred, tgt, B, T, D, loss, dpred, current_stream_ptr(pred.device))
        box['loss'] = loss
        return dpred

    pred, grads = rnn_forward_backward(rnn, x, lengths, dloss, mask)
    return box['loss'], {prefix + k: v for k, v in grads.items()}, pred


@torch.no_grad()
def footcontact_shared_step(module, contact_input, lengths, foot_contacts, mask=None):
    """FootContact.shared_step + backward (footcontact.py:43-65).  contact_input [B,T,132] = cat(noisy target joints, imu) as the
    reference forms it (footcontact.py:57-61; the noise is the caller's, like the dropout mask); foot_contacts [B,T,2] in {0, 1}."""
    def call(lib, pred, tgt, B, T, D, loss, dpred, stream):
        _cabi.check(lib.mp_footcontact_loss(pred.data_ptr(), tgt.data_ptr(), B, T, loss.data_ptr(), dpred.data_ptr(), stream), 'mp_footcontact_loss')
    return _step_with_loss(module.footcontact, 'footcontact.', contact_input, lengths, foot_contacts, mask, call)


@torch.no_grad()
def velocity_shared_step(module, vel_input, lengths, target_vel, mask=None):
    """Velocity.shared_step + backward (velocity.py:50-86).  vel_input [B,T,132] = cat(noisy target joints, imu); target_vel [B,T,72]."""
    def call(lib, pred, tgt, B, T, D, loss, dpred, stream):
        _cabi.check(lib.mp_velocity_loss(pred.data_ptr(), tgt.data_ptr(), B, T, D, loss.data_ptr(), dpred.data_ptr(), stream), 'mp_velocity_loss')
    return _step_with_loss(module.vel, 'vel.', vel_input, lengths, target_vel, mask, call)


@torch.no_grad()
def poser_shared_step(module, pose_input, lengths, target_pose_r6d, target_joints, mask=None):
    """Poser.shared_step + backward (poser.py:65-98).  pose_input [B,T,132] = cat(noisy target joints, imu) (poser.py:81-85; the noise
    is the caller's); target_pose_r6d [B,T,144] the 24 joints' global r6d (the 16 reduced ones are selected here like poser.py:88);
    target_joints [B,T,72]."""
    from .config import joint_set
    lib = _cabi.lib()
    box = {}

    def dloss(pred):
        B, T, D = pred.shape
        pose_t = _f32c(target_pose_r6d.to(pred.device).view(B, T, 24, 6)[:, :, joint_set.reduced].reshape(B, T, 96))
        joints_t = _f32c(target_joints.to(pred.device)).view(B, T, 72)
        loss = torch.zeros((), device=pred.device, dtype=torch.float64)
        dpred = torch.empty_like(pred)
        _cabi.check(lib.mp_poser_loss(pred.data_ptr(), pose_t.data_ptr(), joints_t.data_ptr(), B, T, T_WEIGHT, loss.data_ptr(),
                                      dpred.data_ptr(), current_stream_ptr(pred.device)), 'mp_poser_loss')
        box['loss'] = loss
        return dpred

    pred, grads = rnn_forward_backward(module.pose, pose_input, lengths, dloss, mask)
    return box['loss'], {'pose.' + k: v for k, v in grads.items()}, pred
