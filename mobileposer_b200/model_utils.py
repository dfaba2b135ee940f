"""mobileposer/utils/model_utils.py:6-25 for the B200 path."""
from __future__ import annotations

import torch

from .config import joint_set


def default_device() -> torch.device:
    """model_config.device of the reference (config.py:43), resolved at call time."""
    return torch.device('cuda:0' if torch.cuda.is_available() else 'cpu')


def load_model(model_path: str):
    """Load a MobilePoserNet from a plain state_dict `.pth` (what combine_weights.py:53-56 writes) or, like the
    reference's fallback, from a Lightning checkpoint whose weights sit under 'state_dict'."""
    from .net import MobilePoserNet
    device = default_device()
    model = MobilePoserNet().to(device)
    blob = torch.load(model_path, map_location=device)
    if isinstance(blob, dict) and 'state_dict' in blob and not any(k.startswith('pose.') for k in blob):
        blob = blob['state_dict']
    model.load_state_dict(blob)
    return model


def reduced_pose_to_full(reduced_pose: torch.Tensor) -> torch.Tensor:
    """[B, S, 16*9] rotation matrices of the reduced joints -> [B, S, 24*9] with identities elsewhere
    (model_utils.py:18-25).  Layout helper only; the fused hot path does this inside K5."""
    B, S = reduced_pose.shape[0], reduced_pose.shape[1]
    full = torch.eye(3, device=reduced_pose.device).repeat(B, S, 24, 1, 1)
    full[:, :, joint_set.reduced] = reduced_pose.view(B, S, joint_set.n_reduced, 3, 3)
    return full.view(B, S, -1)


def local6d_to_pose(pose6d: torch.Tensor) -> torch.Tensor:
    """[..., 16, 6] compact local pose (the first two columns of the 16 non-ignored joints' local rotations, what
    `HostOffline(compact=True)` / mp_net_enqueue_offline_host_compact transfer) -> [..., 24, 3, 3] full local pose: third column =
    cross product, ignored joints = identity (net.py:98).  Pure layout + one cross product on the consumer's side; works on any device."""
    lead = pose6d.shape[:-2]
    v = pose6d.reshape(-1, joint_set.n_reduced, 6)
    c0, c1 = v[..., :3], v[..., 3:]
    r = torch.stack((c0, c1, torch.linalg.cross(c0, c1, dim=-1)), dim=-1)          # columns
    full = torch.eye(3, dtype=pose6d.dtype, device=pose6d.device).repeat(v.shape[0], 24, 1, 1)
    full[:, joint_set.reduced] = r
    return full.view(*lead, 24, 3, 3)
