"""Host-side mirror of the reference's four RNN heads (same class names, ctor order, state_dict keys).

    RNN          mobileposer/models/rnn.py:9-33
    Joints       mobileposer/models/joints.py:13-52       RNN(60, 72, 256)
    Poser        mobileposer/models/poser.py:14-63        RNN(132, 96, 256)
    FootContact  mobileposer/models/footcontact.py:13-41  RNN(132, 2, 64)
    Velocity     mobileposer/models/velocity.py:14-48     RNN(132, 72, 256, bidirectional=False), stateful

The torch modules held here (`nn.LSTM`, `nn.Linear`) are *parameter containers only*: they give the
heads the reference's parameter names/shapes/initialisation so a reference `state_dict` loads
unchanged.  Their torch `forward` is never called -- every forward goes through the C ABI
(mobileposer_b200/_cabi.py -> csrc/), and raises if the parameters are not on an sm_100 CUDA device.
Inference semantics only (dropout = identity, no autograd); training is out of scope.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _cabi
from .config import model_config


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f'mobileposer_b200: {what} is on {t.device}; the hot path only runs on a CUDA sm_100 device '
            '(there is no CPU fallback -- move the module and its inputs with .to("cuda")).')


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def current_stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class PackedHead:
    """Owns one `mp_rnn_t` handle; rebuilt when the parameters it was packed from change."""

    def __init__(self):
        self.handle = None
        self.key = None
        self.generation = 0      # bumps on every repack: a freed-and-reallocated handle can come back at the same address

    def get(self, rnn: 'RNN') -> int:
        params = rnn.weight_tensors()
        first = params['linear1.weight']
        _require_cuda(first, 'RNN parameters')
        key = tuple((p.data_ptr(), p._version, p.device.index) for p in params.values())
        if self.handle is not None and key == self.key:
            return self.handle
        self.close()
        lib = _cabi.lib()
        keep = {k: _f32c(v) for k, v in params.items()}
        w = _cabi.RnnWeights()
        w.n_input, w.n_output, w.n_hidden = rnn.n_input, rnn.n_output, rnn.n_hidden
        w.n_layers, w.bidirectional = rnn.n_rnn_layer, int(rnn.bidirectional)
        w.linear1_w, w.linear1_b = keep['linear1.weight'].data_ptr(), keep['linear1.bias'].data_ptr()
        w.linear2_w, w.linear2_b = keep['linear2.weight'].data_ptr(), keep['linear2.bias'].data_ptr()
        for layer in range(min(rnn.n_rnn_layer, 2)):
            for d in range(2 if rnn.bidirectional else 1):
                sfx = f'_l{layer}' + ('_reverse' if d else '')
                w.w_ih[layer][d] = keep['rnn.weight_ih' + sfx].data_ptr()
                w.w_hh[layer][d] = keep['rnn.weight_hh' + sfx].data_ptr()
                w.b_ih[layer][d] = keep['rnn.bias_ih' + sfx].data_ptr()
                w.b_hh[layer][d] = keep['rnn.bias_hh' + sfx].data_ptr()
        out = C.c_void_p()
        with torch.cuda.device(first.device):
            _cabi.check(lib.mp_rnn_create(C.byref(out), C.byref(w), current_stream_ptr(first.device)), 'mp_rnn_create')
        self.handle, self.key = out.value, key
        self.generation += 1
        return self.handle

    def close(self):
        if self.handle is not None:
            _cabi.lib().mp_rnn_destroy(self.handle)
            self.handle = None
            self.key = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RNN(nn.Module):
    """linear1 -> ReLU -> 2-layer LSTM -> linear2 (mobileposer/models/rnn.py:9-33)."""

    def __init__(self, n_input, n_output, n_hidden, n_rnn_layer=2, bidirectional=True, dropout=0.4):
        super().__init__()
        # same construction order as the reference so a seeded default init is identical
        self.rnn = nn.LSTM(input_size=n_hidden, hidden_size=n_hidden, num_layers=n_rnn_layer, bidirectional=bidirectional)
        self.linear1 = nn.Linear(in_features=n_input, out_features=n_hidden)
        self.linear2 = nn.Linear(in_features=n_hidden * (2 if bidirectional else 1), out_features=n_output)
        self.dropout = nn.Dropout(p=dropout)
        self.n_input, self.n_output, self.n_hidden = n_input, n_output, n_hidden
        self.n_rnn_layer, self.bidirectional = n_rnn_layer, bidirectional
        self._packed = PackedHead()

    def weight_tensors(self):
        return {k: v for k, v in self.named_parameters()}

    def packed_handle(self) -> int:
        return self._packed.get(self)

    def packed_key(self):
        """(handle, generation): what caches built on top of the handle (mp_net graphs) must be keyed on."""
        return self._packed.get(self), self._packed.generation

    @torch.no_grad()
    def forward(self, x, seq_lengths=None, h=None, x2=None):
        """Returns (y, output_lengths, (h_n, c_n)) like the reference.

        `x2` (extension): a second input concatenated after `x` on the last dim inside the first GEMM,
        so callers need not materialise torch.cat((joints, imu), -1) (net.py:106,113).
        Without `seq_lengths` the reference's LSTM is sequence-first (rnn.py:15: batch_first=False), i.e.
        dim 0 of `x` is time; that convention is reproduced.
        """
        _require_cuda(x, 'RNN input')
        lib = _cabi.lib()
        handle = self.packed_handle()
        seq_first = seq_lengths is None
        if seq_first:
            x = x.transpose(0, 1)
            x2 = x2.transpose(0, 1) if x2 is not None else None
        xa = _f32c(x)
        xb = _f32c(x2) if x2 is not None else None
        B, T = xa.shape[0], xa.shape[1]
        dev = xa.device
        dirs = 2 if self.bidirectional else 1
        lengths_dev, out_lengths, t_out = None, None, T
        if not seq_first:
            lens = [int(v) for v in (seq_lengths.tolist() if torch.is_tensor(seq_lengths) else seq_lengths)]
            if len(lens) != B:
                raise ValueError(f'seq_lengths has {len(lens)} entries for a batch of {B}')
            if min(lens) <= 0 or max(lens) > T:
                raise RuntimeError('Length of all samples has to be greater than 0 and at most the padded length')
            out_lengths = torch.tensor(lens, dtype=torch.int64)
            t_out = max(lens)
            if min(lens) < T:
                lengths_dev = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
        h0 = c0 = None
        if h is not None:
            h0, c0 = _f32c(h[0]), _f32c(h[1])
            want = (self.n_rnn_layer * dirs, B, self.n_hidden)
            if tuple(h0.shape) != want or tuple(c0.shape) != want:
                raise RuntimeError(f'Expected hidden size {want}, got {tuple(h0.shape)}')
        y = torch.empty(B, T, self.n_output, device=dev, dtype=torch.float32)
        hn = torch.empty(self.n_rnn_layer * dirs, B, self.n_hidden, device=dev, dtype=torch.float32)
        cn = torch.empty_like(hn)
        with torch.cuda.device(dev):
            ws_bytes = lib.mp_rnn_workspace_bytes(handle, B, T)
            ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            _cabi.check(lib.mp_rnn_forward(
                handle, xa.data_ptr(), xa.shape[2], xb.data_ptr() if xb is not None else None,
                xb.shape[2] if xb is not None else 0, B, T,
                lengths_dev.data_ptr() if lengths_dev is not None else None,
                h0.data_ptr() if h0 is not None else None, c0.data_ptr() if c0 is not None else None,
                hn.data_ptr(), cn.data_ptr(), y.data_ptr(), ws.data_ptr(), ws_bytes, current_stream_ptr(dev)),
                'mp_rnn_forward')
        if seq_first:
            y = y.transpose(0, 1)
        elif t_out < T:
            y = y[:, :t_out]
        return y, out_lengths, (hn, cn)


class _Head(nn.Module):
    """Common surface of the four Lightning modules (hypers/finetune kept as plain attributes)."""

    def __init__(self, finetune: bool = False):
        super().__init__()
        self.C = model_config
        self.finetune = finetune


class Joints(_Head):
    def __init__(self, finetune: bool = False):
        super().__init__(finetune)
        self.joints = RNN(self.C.n_imu, 24 * 3, 256)

    def forward(self, batch, input_lengths=None):
        return self.joints(batch, input_lengths)[0]


class Poser(_Head):
    def __init__(self, finetune: bool = False):
        super().__init__(finetune)
        self.pose = RNN(self.C.n_output_joints * 3 + self.C.n_imu, 16 * 6, 256)

    def forward(self, batch, input_lengths=None):
        return self.pose(batch, input_lengths)[0]


class FootContact(_Head):
    def __init__(self):
        super().__init__()
        self.footcontact = RNN(self.C.n_output_joints * 3 + self.C.n_imu, 2, 64)

    def forward(self, batch, input_lengths=None):
        return self.footcontact(batch, input_lengths)[0]


class Velocity(_Head):
    def __init__(self):
        super().__init__()
        self.vel = RNN(self.C.n_output_joints * 3 + self.C.n_imu, 24 * 3, 256, bidirectional=False)
        self.rnn_state = None

    def forward(self, batch, input_lengths=None):
        return self.vel(batch, input_lengths)[0]

    def forward_online(self, batch, input_lengths=None):
        """Stateful: starts from and stores `self.rnn_state` (velocity.py:45-48)."""
        vel, _, self.rnn_state = self.vel(batch, input_lengths, self.rnn_state)
        return vel
