"""Seeded synthetic 5-IMU windows in the reference's 60-float frame layout.

Layout per frame (mobileposer/data.py:69-76): [acc(5x3) / acc_scale | ori(5x3x3)]
with the slots outside the chosen device combo zeroed in BOTH halves.  The
recipe follows SURVEY.md section 8(d): orientations are a random walk on SO(3)
(uniform start, per-frame axis-angle increments ~ N(0, 0.05^2) rad),
accelerations ~ N(0, 3^2) m/s^2 scaled by 1/acc_scale.  One torch CPU generator
per sequence id, so a sequence is the same no matter which rank or batch it
lands in.
"""
from __future__ import annotations

import torch

from .config import amass

BASE_SEED = 1234


def _rodrigues(v: torch.Tensor) -> torch.Tensor:
    """Axis-angle [..., 3] (float64) -> rotation matrices [..., 3, 3]."""
    theta = v.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = v / theta
    kx, ky, kz = k.unbind(-1)
    zero = torch.zeros_like(kx)
    K = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], dim=-1).view(*v.shape[:-1], 3, 3)
    s = torch.sin(theta)[..., None]
    c = torch.cos(theta)[..., None]
    eye = torch.eye(3, dtype=v.dtype).expand_as(K)
    return eye + s * K + (1 - c) * (K @ K)


def _quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y),
    ], dim=-1).view(*q.shape[:-1], 3, 3)


def synthetic_imu_batch(seq_ids, T: int, combo: str = 'lw_rp', base_seed: int = BASE_SEED) -> torch.Tensor:
    """Return [len(seq_ids), T, 60] float32 (CPU)."""
    n = len(seq_ids)
    q0 = torch.empty(n, 5, 4, dtype=torch.float64)
    inc = torch.empty(n, T, 5, 3, dtype=torch.float64)
    acc = torch.empty(n, T, 5, 3, dtype=torch.float64)
    for i, sid in enumerate(seq_ids):
        g = torch.Generator().manual_seed(base_seed + int(sid))
        q0[i] = torch.randn(5, 4, generator=g, dtype=torch.float64)
        inc[i] = torch.randn(T, 5, 3, generator=g, dtype=torch.float64) * 0.05
        acc[i] = torch.randn(T, 5, 3, generator=g, dtype=torch.float64) * 3.0
    dR = _rodrigues(inc)                       # [n, T, 5, 3, 3]
    ori = torch.empty(n, T, 5, 3, 3, dtype=torch.float64)
    cur = _quat_to_matrix(q0)                  # [n, 5, 3, 3]
    for t in range(T):
        cur = cur @ dR[:, t]
        ori[:, t] = cur
    acc = acc / amass.acc_scale
    keep = amass.combos[combo]
    mask = torch.zeros(5, dtype=torch.float64)
    mask[keep] = 1.0
    acc = acc * mask.view(1, 1, 5, 1)
    ori = ori * mask.view(1, 1, 5, 1, 1)
    return torch.cat([acc.flatten(2), ori.flatten(2)], dim=2).float().contiguous()


def synthetic_imu(seq_id: int, T: int, combo: str = 'lw_rp', base_seed: int = BASE_SEED) -> torch.Tensor:
    """Return one [T, 60] float32 window."""
    return synthetic_imu_batch([seq_id], T, combo, base_seed)[0]


def well_conditioned_state_dict(state_dict: dict) -> dict:
    """A seeded state_dict whose pose head emits r6d columns near (1,0,0 | 0,1,0), like a trained model does.

    With the default random init the pose head's r6d columns have norm ~0.05 (down to 0.005 for the orthogonal part),
    so the Gram-Schmidt step of net.py:93-99 amplifies fp32 round-off by up to 200x and the reference's own fp32
    result sits > 1e-4 rad from a float64 evaluation.  Setting `pose.pose.linear2.bias` to [1,0,0,0,1,0] x 16 and
    scaling `pose.pose.linear2.weight` by 0.1 keeps every other tensor of the seeded init and makes the r6d -> rotation
    step well conditioned, so the north star's flat 1e-4 rad can be held on every (frame, joint).  The same dict is
    loaded into the live reference (oracle/make_golden.py, fixtures `wc_*`) and into the CUDA net."""
    sd = {k: v.clone() for k, v in state_dict.items()}
    sd['pose.pose.linear2.bias'] = torch.tensor([1.0, 0, 0, 0, 1.0, 0] * 16, dtype=torch.float32)
    sd['pose.pose.linear2.weight'] = sd['pose.pose.linear2.weight'] * 0.1
    return sd
