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
caller so a step is reproducible); None = eval mode.  `HeadTrainer` is the loop around it (gradient clipping, AdamW, the data-parallel
gradient all-reduce): mp_grad_sq_norm / mp_adamw_step over flat buffers."""
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
def rnn_forward_backward(rnn: RNN, x, lengths, dloss_dy_fn, mask=None, grad_out=None):
    """Forward of one head with saved activations, `dy = dloss_dy_fn(y)` on the device, backward.  -> (y, grads {param name: tensor}).
    grad_out: {param name: preallocated contiguous fp32 tensor} to write the gradients into (HeadTrainer's flat buffer)."""
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
    grads = {k: torch.empty_like(v) for k, v in params.items()} if grad_out is None else grad_out
    for k, v in params.items():
        if tuple(grads[k].shape) != tuple(v.shape) or grads[k].dtype != torch.float32 or not grads[k].is_contiguous():
            raise ValueError(f'grad_out[{k!r}] must be a contiguous float32 tensor of shape {tuple(v.shape)}')
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
def joints_shared_step(module, imu, lengths, target, mask=None, grad_out=None):
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

    pred, grads = rnn_forward_backward(module.joints, imu, lengths, dloss, mask, grad_out)
    return box['loss'], {'joints.' + k: v for k, v in grads.items()}, pred


def _step_with_loss(rnn, prefix, x, lengths, target, mask, loss_call, grad_out=None):
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

    pred, grads = rnn_forward_backward(rnn, x, lengths, dloss, mask, grad_out)
    return box['loss'], {prefix + k: v for k, v in grads.items()}, pred


@torch.no_grad()
def footcontact_shared_step(module, contact_input, lengths, foot_contacts, mask=None, grad_out=None):
    """FootContact.shared_step + backward (footcontact.py:43-65).  contact_input [B,T,132] = cat(noisy target joints, imu) as the
    reference forms it (footcontact.py:57-61; the noise is the caller's, like the dropout mask); foot_contacts [B,T,2] in {0, 1}."""
    def call(lib, pred, tgt, B, T, D, loss, dpred, stream):
        _cabi.check(lib.mp_footcontact_loss(pred.data_ptr(), tgt.data_ptr(), B, T, loss.data_ptr(), dpred.data_ptr(), stream), 'mp_footcontact_loss')
    return _step_with_loss(module.footcontact, 'footcontact.', contact_input, lengths, foot_contacts, mask, call, grad_out)


@torch.no_grad()
def velocity_shared_step(module, vel_input, lengths, target_vel, mask=None, grad_out=None):
    """Velocity.shared_step + backward (velocity.py:50-86).  vel_input [B,T,132] = cat(noisy target joints, imu); target_vel [B,T,72]."""
    def call(lib, pred, tgt, B, T, D, loss, dpred, stream):
        _cabi.check(lib.mp_velocity_loss(pred.data_ptr(), tgt.data_ptr(), B, T, D, loss.data_ptr(), dpred.data_ptr(), stream), 'mp_velocity_loss')
    return _step_with_loss(module.vel, 'vel.', vel_input, lengths, target_vel, mask, call, grad_out)


@torch.no_grad()
def poser_shared_step(module, pose_input, lengths, target_pose_r6d, target_joints, mask=None, grad_out=None):
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

    pred, grads = rnn_forward_backward(module.pose, pose_input, lengths, dloss, mask, grad_out)
    return box['loss'], {'pose.' + k: v for k, v in grads.items()}, pred


# ---- the optimizer loop -------------------------------------------------------------------------------------------------------------
def average_gradients(flat_grads, process_group=None):
    """Data-parallel exchange of a training step: ONE all-reduce (sum) of the head's flat gradient buffer over the ranks (NCCL on the
    device, gloo in the CPU tests).  Returns the factor that turns the sum into the mean (1 / world size) -- it is folded into the
    optimizer kernel's gradient scale, so the mean is never materialised.  No-op (factor 1) without an initialised process group."""
    import torch.distributed as dist
    if process_group is False or not (dist.is_available() and dist.is_initialized()):       # False: explicitly single-process
        return 1.0
    world = dist.get_world_size(process_group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=process_group)
    return 1.0 / world


class HeadTrainer:
    """The training loop of ONE head on the device: what Lightning runs for `overfit.py:41-56` / `train.py:60-98` around the module's
    `training_step` -- zero the gradients, `shared_step`, backward, `clip_grad_norm_(gradient_clip_val)`, the optimizer of
    `configure_optimizers` (joints.py:113-114 and the same line in the other heads: `torch.optim.AdamW(self.parameters(), lr=1e-3)`,
    torch's defaults otherwise).  Lightning itself is not a dependency: the loop is these four calls.

    The head's parameters are moved into ONE flat fp32 buffer (each nn.Parameter becomes a view of it, so `state_dict()` and the
    inference path see the trained values); gradients, exp_avg and exp_avg_sq are flat buffers of the same layout.  A step is the
    head's forward / loss / backward kernels, one all-reduce of the flat gradient buffer when a process group is initialised (data
    parallel: every rank steps on its own shard of the batch), one gradient-norm reduction and one fused clip + AdamW kernel.

        trainer = HeadTrainer(mobileposer_b200.Joints().cuda())
        loss = trainer.training_step(imu, lengths, target_joints, mask=dropout_mask(...))      # the head's shared_step arguments
    """

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, gradient_clip_val=1.0, process_group=None):
        steps = {'joints': joints_shared_step, 'pose': poser_shared_step, 'footcontact': footcontact_shared_step, 'vel': velocity_shared_step}
        found = [(name, fn) for name, fn in steps.items() if isinstance(getattr(module, name, None), RNN)]
        if len(found) != 1:
            raise TypeError('HeadTrainer takes one of the four head modules (Joints, Poser, FootContact, Velocity)')
        self.module, (self.prefix, self.step_fn) = module, found[0]
        self.rnn = getattr(module, self.prefix)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.gradient_clip_val, self.process_group = gradient_clip_val, process_group
        named = list(self.rnn.named_parameters())
        _require_cuda(named[0][1], 'HeadTrainer parameters')
        dev = named[0][1].device
        offsets, total = {}, 0
        for name, p in named:
            offsets[name] = total
            total += (p.numel() + 3) // 4 * 4                   # 16-byte aligned views (the kernels read float4)
        self.flat_params = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grads = torch.zeros_like(self.flat_params)
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.grad_views = {}
        with torch.no_grad():
            for name, p in named:
                o, n = offsets[name], p.numel()
                self.flat_params[o:o + n].copy_(p.detach().reshape(-1).float())
                p.data = self.flat_params[o:o + n].view(p.shape)
                self.grad_views[name] = self.flat_grads[o:o + n].view(p.shape)
        self._sq_norm = torch.zeros((), device=dev, dtype=torch.float64)
        self.global_step = 0
        self.last_grad_norm = None

    @torch.no_grad()
    def training_step(self, *shared_step_args, mask=None):
        """One optimisation step; the positional arguments are the head's `*_shared_step` arguments after the module.
        -> the step's loss (float64 scalar on the device, before the update -- what `training_step` logs, joints.py:77-81)."""
        loss, _, _ = self.step_fn(self.module, *shared_step_args, mask=mask, grad_out=self.grad_views)
        self.optimizer_step()
        return loss

    @torch.no_grad()
    def optimizer_step(self):
        lib = _cabi.lib()
        dev = self.flat_params.device
        scale = average_gradients(self.flat_grads, self.process_group)
        self.global_step += 1
        with torch.cuda.device(dev):
            stream = current_stream_ptr(dev)
            n = self.flat_params.numel()
            sq = None
            if self.gradient_clip_val is not None:
                self._sq_norm.zero_()
                _cabi.check(lib.mp_grad_sq_norm(self.flat_grads.data_ptr(), n, self._sq_norm.data_ptr(), stream), 'mp_grad_sq_norm')
                sq = self._sq_norm.data_ptr()
            _cabi.check(lib.mp_adamw_step(self.flat_params.data_ptr(), self.flat_grads.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), n, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                          self.global_step, sq, float(self.gradient_clip_val or 0.0), float(scale), stream), 'mp_adamw_step')
        # the kernels wrote through raw pointers: tell torch (and the packed inference copies keyed on it) that the values changed
        for p in self.rnn.parameters():
            torch.autograd.graph.increment_version(p)

    def grad_norm(self):
        """Total gradient norm of the last step before clipping (mean over ranks), as clip_grad_norm_ returns it."""
        import torch.distributed as dist
        single = self.process_group is False or not (dist.is_available() and dist.is_initialized())
        return self._sq_norm.sqrt().item() / (1 if single else dist.get_world_size(self.process_group))
