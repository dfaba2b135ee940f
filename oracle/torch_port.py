"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the MobilePoser hot path.

This file is the *checker* for the CUDA product path and the timed CPU baseline
of bench.py (`cpu_baseline`, `--impl reference`).  Nothing in
`mobileposer_b200/` imports it; the product path has no CPU fallback.

It restates, as plain functions over a 72-tensor state_dict, what the
reference computes on CPU (citations relative to /root/reference):

  rnn_head            mobileposer/models/rnn.py:20-33  (linear1 -> ReLU -> [dropout=id in eval]
                      -> pack_padded_sequence -> nn.LSTM(2 layers) -> pad_packed_sequence -> linear2)
  net_forward         mobileposer/models/net.py:101-119
  reduced_global_to_full  net.py:93-99 + articulate/math/angular.py:167-182
                      + utils/model_utils.py:18-25 + articulate/math/spatial.py:115-123,197-221
  offline_translation net.py:121-154   (K6)
  OnlineState.step    net.py:173-219   (K7) + reset net.py:84-88

The LSTM/Linear arithmetic is a third-party dependency of the reference
(torch; pinned torch==2.1.2 in requirements.txt:121, this image has
2.11.0+cu128).  Like the reference's call site (rnn.py:15,27) the port calls
torch's own CPU nn.LSTM, so it times and computes what the reference would on
these host cores.  Parity pinning: the reference ships no tests or golden
vectors; this port is pinned against the *live reference
run in the build container* by oracle/make_golden.py, whose outputs are the
fixtures in tests/golden/ (tests/test_oracle.py re-checks the port against
them on every run).  The physics hook (net.py:157-169) has no source in the
reference tree: parity unpinned for that row, see DESIGN.md.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from mobileposer_b200.config import (HEAD_PREFIX, HEAD_SHAPES, FLOOR_Y, SMPL_J_ZERO, SMPL_PARENT,
                                     PROB_THRESHOLD, amass, datasets, joint_set, model_config)


class _Head:
    """One RNN head rebuilt from state_dict tensors (rnn.py:13-18)."""

    def __init__(self, sd, prefix, n_in, n_out, hidden, bidir, dtype=torch.float32):
        self.w1, self.b1 = sd[prefix + 'linear1.weight'], sd[prefix + 'linear1.bias']
        self.w2, self.b2 = sd[prefix + 'linear2.weight'], sd[prefix + 'linear2.bias']
        self.lstm = torch.nn.LSTM(hidden, hidden, num_layers=2, bidirectional=bidir).to(dtype)
        self.lstm.load_state_dict({k[len(prefix) + 4:]: v for k, v in sd.items()
                                   if k.startswith(prefix + 'rnn.')})
        self.lstm.eval()
        for p in self.lstm.parameters():
            p.requires_grad_(False)

    def __call__(self, x, lengths, state=None):
        a = torch.relu(F.linear(x, self.w1, self.b1))
        packed = pack_padded_sequence(a, lengths, batch_first=True, enforce_sorted=False)
        out, state = self.lstm(packed, state)
        out, _ = pad_packed_sequence(out, batch_first=True)
        return F.linear(out, self.w2, self.b2), state


class OraclePoser:
    """CPU port of MobilePoserNet's inference surface for a given state_dict.

    dtype=torch.float64 evaluates the same torch kernels in double precision (inputs must be double too): the arbiter
    between two fp32 implementations at sizes where the numpy statement (np_port.py) is too slow.  Only `forward` and
    the heads are meant to be used in that mode."""

    def __init__(self, state_dict, dtype=torch.float32):
        sd = {k: v.detach().to(dtype).cpu() for k, v in state_dict.items()}
        self.dtype = dtype
        self.heads = {name: _Head(sd, HEAD_PREFIX[name], *HEAD_SHAPES[name], dtype=dtype) for name in HEAD_SHAPES}
        self.j = torch.tensor(SMPL_J_ZERO, dtype=torch.float32)
        self.floor_y = FLOOR_Y
        self.vel_state = None           # Velocity.rnn_state (velocity.py:30,45-48)
        self.reset_online()

    # ---- heads ---------------------------------------------------------------------------
    def joints_head(self, imu, lengths):
        return self.heads['joints'](imu, lengths)[0]

    def forward(self, imu, lengths):
        """net.py:101-119; returns (pose [B*T,24,3,3], joints [B,T,72], vel, contact [B,T,2])."""
        joints = self.heads['joints'](imu, lengths)[0]
        feat = torch.cat((joints, imu), dim=-1)
        r6d = self.heads['pose'](feat, lengths)[0]
        pose = reduced_global_to_full(r6d)
        contact = self.heads['foot_contact'](feat, lengths)[0]
        vel, self.vel_state = self.heads['velocity'](feat, lengths, self.vel_state)
        return pose, joints, vel.squeeze(0), contact

    # ---- offline (K6) ----------------------------------------------------------------------
    def forward_offline(self, imu, lengths):
        """net.py:121-154 for one sequence (B = 1)."""
        pose, joints, vel, contact = self.forward(imu, lengths)
        tran = offline_translation(joints.squeeze(0), vel, contact.squeeze(0), self.floor_y)
        return pose, joints, tran, contact.squeeze(0)

    # ---- online (K7) -----------------------------------------------------------------------
    def reset_online(self):
        """ctor values net.py:59-64; `reset` only clears imu/root (net.py:84-88)."""
        self.last_lfoot, self.last_rfoot = self.j[10].clone(), self.j[11].clone()
        self.reset()

    def reset(self):
        self.imu = None
        self.current_root_y = 0.0
        self.last_root_pos = torch.zeros(3)

    def forward_online(self, frame):
        """net.py:173-219 for one 60-float frame."""
        W, P = model_config.total_frames, model_config.past_frames
        imu = frame.repeat(W, 1) if self.imu is None else torch.cat((self.imu[1:], frame.view(1, -1)))
        pose, joints, vel, contact = self.forward(imu.unsqueeze(0), [W])
        pose = pose[P].view(-1, 9)
        jp = joints.squeeze(0)[P].view(24, 3)
        c = contact[0][P]
        lf, rf = jp[10], jp[11]
        g = torch.tensor([0, joint_set.gravity_velocity, 0])
        cvel = (self.last_lfoot - lf + g) if c[0] > c[1] else (self.last_rfoot - rf + g)
        pvel = vel.view(-1, 24, 3)[:, 0][P] / (datasets.fps / amass.vel_scale)
        w = prob_to_weight(c.max())                      # NB: no sigmoid online (net.py:197)
        v = pvel * (1 - w) + cvel * w
        foot_y = self.current_root_y + min(lf[1].item(), rf[1].item())
        if foot_y + v[1].item() <= self.floor_y:
            v[1] = self.floor_y - foot_y
        self.current_root_y += v[1].item()
        self.last_lfoot, self.last_rfoot = lf, rf
        self.imu = imu
        self.last_root_pos = self.last_root_pos + v
        return pose, joints.squeeze(0), self.last_root_pos.clone(), c


def prob_to_weight(p):
    lo, hi = PROB_THRESHOLD
    return (p.clamp(lo, hi) - lo) / (hi - lo)


def r6d_to_matrix(r6d):
    """articulate/math/angular.py:167-182 (columns c0, c1, c2 stacked on the last dim; NaN -> 0)."""
    v = r6d.reshape(-1, 6)
    a, b = v[:, :3], v[:, 3:]
    c0 = a / a.norm(dim=1, keepdim=True)
    b = b - (c0 * b).sum(dim=1, keepdim=True) * c0
    c1 = b / b.norm(dim=1, keepdim=True)
    c2 = torch.linalg.cross(c0, c1, dim=1)
    r = torch.stack((c0, c1, c2), dim=-1)
    return torch.where(torch.isnan(r), torch.zeros_like(r), r)


def reduced_global_to_full(r6d):
    """net.py:93-99: r6d [*, 96] -> local rotations [N, 24, 3, 3]."""
    glb16 = r6d_to_matrix(r6d).view(-1, joint_set.n_reduced, 3, 3)
    n = glb16.shape[0]
    glb = torch.eye(3, dtype=r6d.dtype).repeat(n, 24, 1, 1)
    glb[:, joint_set.reduced] = glb16
    loc = torch.empty_like(glb)
    loc[:, 0] = glb[:, 0]
    for i in range(1, 24):
        loc[:, i] = glb[:, SMPL_PARENT[i]].transpose(1, 2) @ glb[:, i]
    loc[:, joint_set.ignored] = torch.eye(3, dtype=r6d.dtype)
    loc[:, 0] = glb[:, 0]
    return loc


def offline_translation(joints72, vel72, contact, floor_y=FLOOR_Y):
    """net.py:129-154 on one sequence: joints72 [T,72], vel72 [T,72], contact [T,2] -> tran [T,3]."""
    jp = joints72.view(-1, 24, 3)
    T = jp.shape[0]
    g = torch.tensor([0, joint_set.gravity_velocity, 0])
    z = torch.zeros(1, 3)
    dl = torch.cat((z, jp[:-1, 10] - jp[1:, 10]))
    dr = torch.cat((z, jp[:-1, 11] - jp[1:, 11]))
    idx = contact.max(dim=1).indices.view(-1, 1)
    cvel = g + (dl * (1 - idx) + dr * idx)
    pvel = vel72.view(-1, 24, 3)[:, 0] / (datasets.fps / amass.vel_scale)
    w = prob_to_weight(contact.max(dim=1).values.sigmoid()).view(-1, 1)
    v = pvel * (1 - w) + cvel * w
    cur = 0.0
    for i in range(T):
        foot = cur + jp[i, 10:12, 1].min().item()
        if foot + v[i, 1].item() <= floor_y:
            v[i, 1] = floor_y - foot
        cur += v[i, 1].item()
    return torch.stack([v[:i + 1].sum(dim=0) for i in range(T)])


def offline_batched(oracle: OraclePoser, imu, lengths, reset_velocity=True):
    """Per-sequence semantics of a batch: B independent forward_offline calls.

    Returns lists of per-sequence (pose [L,24,3,3], joints [L,72], tran [L,3], contact [L,2]).
    """
    outs = []
    for b, L in enumerate(lengths):
        if reset_velocity:
            oracle.vel_state = None
        pose, joints, tran, contact = oracle.forward_offline(imu[b:b + 1, :L], [L])
        outs.append((pose, joints.squeeze(0), tran, contact))
    return outs
